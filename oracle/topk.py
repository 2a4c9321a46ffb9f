"""Score contraction + top-k with a canonical tie rule.  TEST INFRASTRUCTURE.

Reference lines: ``scores = torch.mm(user_emb, known_post_emb.T)`` inference.py:427 and
``torch.topk(scores, min(K, len(scores)))`` inference.py:428 (also train_gnn.py:335-341,494-499,
test_gnn.py:224-231).  ``torch.topk`` breaks ties arbitrarily on CPU, and post-ReLU embeddings
give exact-zero scores, so ids are compared under the total order (score desc, id asc); the
literal ``torch.topk`` is used for VALUES always and for ids only when the K+1 largest scores
are pairwise distinct.
"""
from __future__ import annotations

import torch


def topk_canonical(scores: torch.Tensor, k: int, id_offset: int = 0):
    """Row-wise top-min(k, n) under (score desc, id asc).  A stable descending sort of an
    id-ascending array yields id-ascending order inside ties."""
    n = scores.size(-1)
    k = min(k, n)
    vals, idx = torch.sort(scores, dim=-1, descending=True, stable=True)
    return vals[..., :k].contiguous(), (idx[..., :k] + id_offset).contiguous()


def score_topk(q: torch.Tensor, cat: torch.Tensor, k: int, id_offset: int = 0):
    """inference.py:427-428 for a batch of queries ``q[B,H]`` against ``cat[P,H]``."""
    scores = torch.mm(q, cat.T)
    return topk_canonical(scores, k, id_offset)


def merge_topk(vals_list, ids_list, k: int):
    """Merge per-shard (vals, global ids) lists under the canonical order: identical to the
    unsharded result."""
    vals = torch.cat(vals_list, dim=-1)
    ids = torch.cat(ids_list, dim=-1)
    # order by (val desc, id asc): sort ids ascending first, then a stable sort on values
    o1 = torch.argsort(ids, dim=-1, stable=True)
    vals, ids = vals.gather(-1, o1), ids.gather(-1, o1)
    o2 = torch.argsort(vals, dim=-1, descending=True, stable=True)
    k = min(k, vals.size(-1))
    return vals.gather(-1, o2)[..., :k].contiguous(), ids.gather(-1, o2)[..., :k].contiguous()
