"""Data formats either side of K0: the reference's graph artefact, vectorised
edge-list construction, and an on-disk cache of the destination-sorted CSR structures.

* ``build_edge_index`` -- train_gnn.py:40-73 (``build_edge_index_safe``) without the ``iterrows`` loop.
* ``interaction_type_table`` -- train_gnn.py:226-237 without the two python loops.
* ``temporal_split`` -- the chronological 80/10/10 split of train_gnn.py:28-35.
* ``user_structural_features`` -- build_graph.py:409-429 (social in/out degree and engagement count per
  user, log(1 + x)) without the ``iterrows`` loop.
* ``hetero_inputs`` -- the ``HeteroData`` assembly of train_gnn.py:115-142 as plain dicts.
* ``save_csr_cache`` / ``load_csr_cache`` -- K0 output persisted next to
  ``synthetic_processed_with_semantics.pt`` (build_graph.py:476-486) so a restart skips the sort.
"""
from __future__ import annotations

import torch

from .graph import CSR, RelationGraph, relation_graph
from .nn import REL_DIRECT, REL_ENGAGE, REL_SOCIAL


def build_edge_index(df, user_to_idx, post_to_idx):
    """Vectorised ``build_edge_index_safe`` (train_gnn.py:40-73): map the ``engager`` /
    ``target_user`` / ``post_id`` columns through the id dictionaries, keep rows where all three map
    (same row order), return ``(engage_edge[2,E], author_edge[2,E])`` with GLOBAL post ids."""
    import pandas as pd
    eng = df["engager"].map(user_to_idx)
    tgt = df["target_user"].map(user_to_idx)
    post = df["post_id"].map(post_to_idx)
    ok = eng.notna() & tgt.notna() & post.notna()
    to_t = lambda s: torch.as_tensor(pd.to_numeric(s[ok]).to_numpy(dtype="int64"), dtype=torch.long)
    engager, post_global, target_user = to_t(eng), to_t(post), to_t(tgt)
    return torch.stack([engager, post_global]), torch.stack([post_global, target_user])


def interaction_type_table(train_interactions, post_to_idx, device=None):
    """Vectorised train_gnn.py:226-237: ``w[global post id] = 3.0`` for "QT" rows, ``1.0`` otherwise,
    ``0.0`` for ids never seen in training; when a post id repeats the LAST row wins (dict overwrite)."""
    size = max(post_to_idx.values()) + 1
    gid = train_interactions["post_id"].map(post_to_idx).to_numpy(dtype="int64")
    w = torch.where(torch.as_tensor((train_interactions["interaction"] == "QT").to_numpy().copy()),
                    torch.tensor(3.0), torch.tensor(1.0))
    table = torch.zeros(size, dtype=torch.float32)
    table[torch.as_tensor(gid)] = w          # duplicate indices: last write wins, like the dict
    # index_put with duplicates is only ordered on CPU; enforce "last wins" explicitly
    order = torch.arange(gid.size)
    last = torch.zeros(size, dtype=torch.long).scatter_reduce_(0, torch.as_tensor(gid), order, "amax", include_self=True)
    seen = torch.zeros(size, dtype=torch.bool)
    seen[torch.as_tensor(gid)] = True
    table[seen] = w[last[seen]]
    return table.to(device) if device is not None else table


def temporal_split(activity, train_frac=0.8, val_frac=0.1):
    """train_gnn.py:28-35: stable chronological sort on ``timestamp``, then ``int(0.8 n)`` / ``int(0.9 n)``
    cut points; returns ``(train, val, test)`` frames (``iloc`` slices of the sorted frame, index reset)."""
    activity_sorted = activity.sort_values("timestamp").reset_index(drop=True)
    n = len(activity_sorted)
    train_end, val_end = int(train_frac * n), int((train_frac + val_frac) * n)
    return activity_sorted.iloc[:train_end], activity_sorted.iloc[train_end:val_end], activity_sorted.iloc[val_end:]


def user_structural_features(social_mapped, activity, user_to_idx, num_users):
    """build_graph.py:409-429 vectorised: ``log(1 + [in-degree, out-degree, engagement count])`` per user as
    fp32 ``[U, 3]``.  ``social_mapped`` holds integer ``follower`` / ``followee`` node ids (one row per edge,
    duplicates counted); the engagement count is ``value_counts`` of the mapped ``engager`` column (unmapped
    engagers dropped, like ``map`` + ``value_counts`` does)."""
    import numpy as np
    f = social_mapped["follower"].to_numpy(dtype="int64")
    t = social_mapped["followee"].to_numpy(dtype="int64")
    out_social = np.bincount(f, minlength=num_users).astype(np.float64)
    in_social = np.bincount(t, minlength=num_users).astype(np.float64)
    eng = activity["engager"].map(user_to_idx).dropna().to_numpy(dtype="int64")
    engagement = np.bincount(eng, minlength=num_users).astype(np.float64)
    feats = np.stack([np.log(in_social + 1), np.log(out_social + 1), np.log(engagement + 1)], axis=1)
    return torch.tensor(feats, dtype=torch.float)


def _edge_checksum(ei):
    """Order-sensitive 64-bit fingerprint of an ``edge_index`` tensor (wrapping int64 arithmetic): equal
    counts are not enough to trust a cached CSR -- another split of the same size must not install it."""
    e = ei.size(1)
    if e == 0:
        return 0
    pos = torch.arange(1, e + 1, device=ei.device, dtype=torch.int64)
    mix = (ei[0] * 1000003 + ei[1]) * (2 * pos + 1)
    return int(mix.sum().item())


def hetero_inputs(data, train_engage_edges, device=None):
    """``x_dict`` / ``edge_index_dict`` exactly as train_gnn.py:115-142 builds them on ``HeteroData``:
    features split at ``num_users``; social edges as stored; engagement edges from the TRAINING period
    with post ids shifted to local (``- num_users``) and range-masked (:128-133); reverse edges by
    ``.flip(0)`` (:142)."""
    nu, np_ = int(data["num_users"]), int(data["num_posts"])
    x = data["x"]
    src, dst = data["edge_index_social"]
    social = torch.stack([src, dst], dim=0)
    post_local = train_engage_edges[1] - nu
    mask = (train_engage_edges[0] < nu) & (post_local >= 0) & (post_local < np_)
    engages = torch.stack([train_engage_edges[0][mask], post_local[mask]], dim=0)
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    x_dict = {"user": mv(x[:nu].contiguous()), "post": mv(x[nu:].contiguous())}
    ei = {REL_SOCIAL: mv(social.contiguous()), REL_ENGAGE: mv(engages.contiguous()),
          REL_DIRECT: mv(engages.flip(0).contiguous())}
    return x_dict, ei


def save_csr_cache(path, edge_index_dict, x_dict):
    """Persist the K0 structures (forward CSR, and the transpose when already built) of every relation."""
    out = {}
    for rel, ei in edge_index_dict.items():
        g = relation_graph(ei, x_dict[rel[0]].size(0), x_dict[rel[2]].size(0))
        ent = {"n_src": g.n_src, "n_dst": g.n_dst, "n_edges": int(ei.size(1)), "checksum": _edge_checksum(ei),
               "fwd": {k: getattr(g.fwd, k).cpu() for k in ("rowptr", "col", "eid")}}
        if g._bwd is not None:
            ent["bwd"] = {k: getattr(g._bwd, k).cpu() for k in ("rowptr", "col", "eid")}
        out["|".join(rel)] = ent
    torch.save(out, path)


def load_csr_cache(path, edge_index_dict, x_dict):
    """Install cached CSR structures for these ``edge_index`` tensors (no K0 launch).  Shapes AND an
    order-sensitive checksum of the edge list are checked against the tensors; a mismatching entry is
    ignored (it will be rebuilt on first use)."""
    cache = torch.load(path, weights_only=True)
    hits = 0
    for rel, ei in edge_index_dict.items():
        ent = cache.get("|".join(rel))
        n_src, n_dst = x_dict[rel[0]].size(0), x_dict[rel[2]].size(0)
        if not ent or ent["n_src"] != n_src or ent["n_dst"] != n_dst or ent["n_edges"] != int(ei.size(1)):
            continue
        if ent.get("checksum") != _edge_checksum(ei):
            continue
        g = relation_graph(ei, n_src, n_dst)
        mk = lambda d, rows, cols: CSR(d["rowptr"].to(ei.device), d["col"].to(ei.device), d["eid"].to(ei.device), rows, cols)
        g._fwd = mk(ent["fwd"], n_dst, n_src)
        if "bwd" in ent:
            g._bwd = mk(ent["bwd"], n_src, n_dst)
        hits += 1
    return hits
