"""Destination-sorted CSR structures built on the GPU from the reference's COO ``edge_index``.

Input format (unchanged from the reference): ``edge_index[2, E]`` int64, row 0 = source ids,
row 1 = destination ids, local per node type, unsorted, duplicates kept
(build_graph.py:387,394,402; train_gnn.py:128-133,142).  The CSR is an internal cache: the same
``edge_index`` tensors are passed every epoch (train_gnn.py:254), so each relation is sorted once.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch

from . import _lib


#: rows with more edges than this are split into slices summed by a second kernel (load balance on
#: skewed / power-law degree distributions; SURVEY.md §7 hard part 2)
LONG_ROW_THRESHOLD = 1024


@dataclass
class LongRows:
    """Virtual-row view of a CSR with long rows cut into <= T-edge slices (trg_long_rows)."""
    vrowptr: torch.Tensor    # int32 [n_vrows + 1]
    vinfo: torch.Tensor      # int32 [n_vrows]: row id, or -(slot + 1) for a slice of a long row
    long_rows: torch.Tensor  # int32 [n_long]
    long_ptr: torch.Tensor   # int32 [n_long + 1]
    n_slots: int


@dataclass
class CSR:
    """Rows = keys (destinations for the forward structure). All int32 on the device."""
    rowptr: torch.Tensor   # [n_rows + 1]
    col: torch.Tensor      # [E]  the other endpoint of each edge, in stable key order
    eid: torch.Tensor      # [E]  original edge position (== argsort(key, stable))
    n_rows: int
    n_cols: int
    _long: object = None   # LongRows | False (no long rows) | None (not computed yet)

    @property
    def n_edges(self) -> int:
        return int(self.col.numel())

    def long_rows(self, threshold: int = None):
        """``LongRows`` when some row exceeds ``threshold`` edges, else ``None``.  Computed once per
        CSR (one host read of the maximum degree); static graphs only pay this at cache-fill time."""
        if self._long is None:
            self._long = _split_long_rows(self, LONG_ROW_THRESHOLD if threshold is None else threshold) or False
        return self._long or None


def _split_long_rows(csr: "CSR", t: int):
    if csr.n_rows == 0 or csr.n_edges <= t:
        return None
    rp = csr.rowptr.long()
    deg = rp[1:] - rp[:-1]
    if int(deg.max()) <= t:
        return None
    is_long = deg > t
    counts = torch.where(is_long, (deg + t - 1) // t, torch.ones_like(deg))
    n_vrows = int(counts.sum())
    orig = torch.repeat_interleave(torch.arange(csr.n_rows, device=rp.device), counts)
    vstart = torch.cumsum(counts, 0) - counts                      # first virtual row of each row
    k = torch.arange(n_vrows, device=rp.device) - vstart[orig]     # slice index inside its row
    vrowptr = torch.empty(n_vrows + 1, dtype=torch.int32, device=rp.device)
    vrowptr[:-1] = (rp[:-1][orig] + k * t).int()
    vrowptr[-1] = csr.n_edges
    part = is_long[orig]
    slot = torch.cumsum(part.long(), 0) - 1
    vinfo = torch.where(part, -(slot + 1), orig).int()
    long_ids = is_long.nonzero().flatten()
    lp = torch.zeros(long_ids.numel() + 1, dtype=torch.long, device=rp.device)
    lp[1:] = torch.cumsum(counts[long_ids], 0)
    return LongRows(vrowptr, vinfo, long_ids.int(), lp.int(), int(lp[-1]))


def build_csr(other: torch.Tensor, key: torch.Tensor, n_key: int, n_other: int,
              validate: bool = True, per_step: bool = False) -> CSR:
    """``trg_csr_build``: stable sort of the edges by ``key``.  Bit-exact with
    ``argsort(key, stable)`` / ``bincount`` / ``cumsum`` (oracle/csr.py).

    ``per_step``: the structure lives for one training step (this step's sampled negatives,
    train_gnn.py:272).  Its rows are not inspected for long-row splitting: that check reads the maximum
    degree on the host, and a host sync inside every step stops the CPU from running ahead of the GPU
    (measured: the whole step's launch overhead then lands on the critical path at 4-8 GPUs).
    Splitting is a load-balance measure only -- results are identical without it."""
    lib = _lib.load()
    if key.dtype != torch.int64 or other.dtype != torch.int64:
        raise TypeError("edge_index must be int64 (torch.long), as in the reference")
    key = key.contiguous()
    other = other.contiguous()
    e = int(key.numel())
    dev = key.device
    if validate and e > 0:
        # one host sync per relation, at cache-fill time only (PyG raises similarly on bad ids)
        lo_k, hi_k = int(key.min()), int(key.max())
        lo_o, hi_o = int(other.min()), int(other.max())
        if lo_k < 0 or hi_k >= n_key or lo_o < 0 or hi_o >= n_other:
            raise IndexError(f"edge_index out of range: key in [{lo_k},{hi_k}] vs {n_key}, "
                             f"other in [{lo_o},{hi_o}] vs {n_other}")
    rowptr = torch.empty(n_key + 1, dtype=torch.int32, device=dev)
    col = torch.empty(e, dtype=torch.int32, device=dev)
    eid = torch.empty(e, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.trg_csr_workspace_bytes(e, n_key))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("trg_csr_build", e * 80 + 8 * (n_key + 1), lib.trg_csr_build,
              _lib.ptr(other) if e else None, _lib.ptr(key) if e else None, e, n_key, _lib.ptr(rowptr),
              _lib.ptr(col) if e else None, _lib.ptr(eid) if e else None, _lib.ptr(ws), ws_bytes,
              _lib.stream())
    return CSR(rowptr, col, eid, n_key, n_other, False if per_step else None)


class RelationGraph:
    """One relation's forward CSR (rows = destinations) and, lazily, its transpose (rows =
    sources) for the atomic-free backward."""

    def __init__(self, edge_index: torch.Tensor, n_src: int, n_dst: int):
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must have shape [2, E]")
        self.edge_index = edge_index  # keep alive: the cache key uses its data_ptr
        self.n_src, self.n_dst = int(n_src), int(n_dst)
        self._fwd = None
        self._bwd = None
        self.inv_deg = None   # fp32 [n_dst], 1 / max(in-degree, 1); filled by the first aggregation

    @property
    def fwd(self) -> CSR:
        if self._fwd is None:
            ei = self.edge_index
            self._fwd = build_csr(ei[0], ei[1], self.n_dst, self.n_src)
        return self._fwd

    @property
    def bwd(self) -> CSR:
        if self._bwd is None:
            ei = self.edge_index
            self._bwd = build_csr(ei[1], ei[0], self.n_src, self.n_dst, validate=self._fwd is None)
        return self._bwd


class PushRelation:
    """A relation partitioned by SOURCE on multi-GPU runs (see collectives.PushMeanAggFn): ``rel``
    holds this rank's edges as (local source id, global destination id); ``inv_deg`` is 1 / max(global
    in-degree, 1) of the destination rows this rank owns."""

    def __init__(self, rel: RelationGraph, inv_deg: torch.Tensor):
        self.rel = rel
        self.inv_deg = inv_deg


class GraphCache:
    """LRU cache keyed on ``(data_ptr, shape, _version, n_src, n_dst)``."""

    def __init__(self, capacity: int = 32):
        self.capacity = capacity
        self._d: OrderedDict = OrderedDict()

    def get(self, edge_index: torch.Tensor, n_src: int, n_dst: int) -> RelationGraph:
        key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version,
               int(n_src), int(n_dst), str(edge_index.device))
        g = self._d.get(key)
        if g is None:
            g = RelationGraph(edge_index, n_src, n_dst)
            self._d[key] = g
            while len(self._d) > self.capacity:
                self._d.popitem(last=False)
        else:
            self._d.move_to_end(key)
        return g

    def clear(self):
        self._d.clear()


_GLOBAL_CACHE = GraphCache()


def relation_graph(edge_index, n_src, n_dst) -> RelationGraph:
    return _GLOBAL_CACHE.get(edge_index, n_src, n_dst)


def clear_cache():
    _GLOBAL_CACHE.clear()
