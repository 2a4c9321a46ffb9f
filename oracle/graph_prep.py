"""Literal restatement of the reference's python-loop graph preparation.  TEST INFRASTRUCTURE.
``build_edge_index_safe`` = train_gnn.py:40-73; ``interaction_type_table`` = train_gnn.py:226-237."""
import torch


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
