"""Bit-exact contract for the destination-sorted CSR (and its transpose).  TEST INFRASTRUCTURE.

Input format = the reference's COO ``edge_index[2, E]`` int64 (row 0 = src, row 1 = dst, local
ids per node type, unsorted, duplicates kept): build_graph.py:387,394,402; local-id shift
train_gnn.py:128-133; reverse edges by ``.flip(0)`` train_gnn.py:142.
"""
from __future__ import annotations

import torch


def csr_by_dst(edge_index: torch.Tensor, n_dst: int):
    """``perm = argsort(dst, stable)``; ``rowptr = [0, cumsum(bincount(dst))]``; ``col = src[perm]``;
    ``eid = perm``.  Duplicates kept, no coalescing, no self loops added."""
    src, dst = edge_index[0].long().cpu(), edge_index[1].long().cpu()
    perm = torch.argsort(dst, stable=True)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long),
                        torch.cumsum(torch.bincount(dst, minlength=n_dst), 0)])
    return rowptr, src[perm], perm


def csr_by_src(edge_index: torch.Tensor, n_src: int):
    """The transposed structure: same with the roles of src and dst swapped."""
    return csr_by_dst(edge_index.flip(0), n_src)
