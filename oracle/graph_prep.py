"""Literal restatement of the reference's python-loop graph preparation.  TEST INFRASTRUCTURE.
``build_edge_index_safe`` = train_gnn.py:40-73; ``interaction_type_table`` = train_gnn.py:226-237;
``temporal_split`` = train_gnn.py:28-35; ``user_structural_features`` = build_graph.py:409-429."""
import numpy as np
import torch


def temporal_split(activity_sub):
    activity_sorted = activity_sub.sort_values("timestamp").reset_index(drop=True)
    n = len(activity_sorted)
    train_end = int(0.8 * n)
    val_end = int(0.9 * n)
    return activity_sorted.iloc[:train_end], activity_sorted.iloc[train_end:val_end], activity_sorted.iloc[val_end:]


def user_structural_features(social_mapped, activity_sub, user_to_idx, U):
    in_social = np.zeros(U)
    out_social = np.zeros(U)
    engagement_count = np.zeros(U)
    for _, row in social_mapped.iterrows():
        f, t = int(row["follower"]), int(row["followee"])
        out_social[f] += 1
        in_social[t] += 1
    eng_counts = activity_sub["engager"].map(user_to_idx).value_counts()
    for uid, cnt in eng_counts.items():
        engagement_count[int(uid)] = cnt
    return torch.tensor(np.stack([np.log(in_social + 1), np.log(out_social + 1), np.log(engagement_count + 1)],
                                 axis=1), dtype=torch.float)


def build_edge_index_safe(df, user_to_idx, post_to_idx):
    engager, post_global, target_user = [], [], []
    for _, row in df.iterrows():
        u_eng = user_to_idx.get(row["engager"])
        u_tgt = user_to_idx.get(row["target_user"])
        p_global = post_to_idx.get(row["post_id"])
        if u_eng is not None and u_tgt is not None and p_global is not None:
            engager.append(u_eng)
            post_global.append(p_global)
            target_user.append(u_tgt)
    engager = torch.tensor(engager, dtype=torch.long)
    post_global = torch.tensor(post_global, dtype=torch.long)
    target_user = torch.tensor(target_user, dtype=torch.long)
    return torch.stack([engager, post_global], dim=0), torch.stack([post_global, target_user], dim=0)


def interaction_type_table(train_interactions, post_to_idx):
    global_post_to_interaction = {}
    for _, row in train_interactions.iterrows():
        global_post_to_interaction[post_to_idx[row["post_id"]]] = row["interaction"]
    t = torch.zeros(max(post_to_idx.values()) + 1, dtype=torch.float32)
    for gid, inter in global_post_to_interaction.items():
        t[gid] = 3.0 if inter == "QT" else 1.0
    return t
