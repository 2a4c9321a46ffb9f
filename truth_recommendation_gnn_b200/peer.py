"""Collectives of the sharded train step over PEER MEMORY (NVLink / NVSwitch), SURVEY.md §8e last row.

The step's two bandwidth collectives are not library calls here:

* reduce-scatter of a partial-sum table + the row finish that followed it = ONE kernel
  (``trg_peer_reduce_rows``): every rank produces its full-height partial table straight into a symmetric
  buffer its peers have mapped; the owner of a row range reads that range from all G buffers with plain
  loads (its own from HBM, G - 1 over NVLink), adds them in rank order and applies 1/deg, the local
  gradient term and the ReLU backward in the same pass.  The reduced table is never materialised.
* all-gather of owned rows = copy-engine pulls: the owned rows are staged into a symmetric buffer and every
  rank copies its peers' chunks with ``cudaMemcpyAsync`` (DMA engines over NVLink): no SM-resident kernel
  competes with the aggregation kernel that runs meanwhile.

Ordering across ranks uses the symmetric-memory signal pads (``barrier`` on the communication stream); the
buffers rotate (two of each), which makes one barrier per collective enough -- see ``PeerComm``.

``torch.distributed._symmetric_memory`` supplies allocation, the rendezvous (CUDA VMM handles exchanged
through the process group's store) and the barrier; nothing here touches the data path of NCCL.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


class _Done:
    """Result of an asynchronous peer collective: ``wait()`` orders the current stream after it."""

    def __init__(self, event, out, keep=None):
        self.event, self.out, self.keep = event, out, keep

    def wait(self):
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None
        self.keep = None
        return self.out


class PeerComm:
    """Symmetric buffers + one communication stream of this rank.

    ``part_elems``: elements of the largest partial table (``[G * rows_per_rank, feat]``), ``part_dtype`` its
    element type; ``stage_elems`` / ``stage_dtype``: the largest owned-row block that is all-gathered.

    Hazards and why one barrier per collective is enough.  All peer collectives of a step are issued in the
    same order on every rank and run in that order on each rank's communication stream: ``[barrier_k,
    transfer_k]``.  A rank reaches ``barrier_{k+1}`` only after its own ``transfer_k`` has finished, so once a
    rank is through ``barrier_{k+1}`` NO peer still reads what transfer ``k`` read.  With two buffers used
    alternately, the buffer a producer writes for collective ``k+2`` was last read by transfer ``k``; the
    producer is ordered after this rank's ``barrier_{k+1}`` (``acquire_partial`` waits for that event on the
    compute stream; the staging copy of an all-gather runs on the communication stream itself)."""

    N_BUF = 2

    def __init__(self, part_elems: int, part_dtype, stage_elems: int, stage_dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device
        self.part_dtype, self.stage_dtype = part_dtype, stage_dtype
        self.part_elems, self.stage_elems = int(part_elems), int(stage_elems)
        self._part, self._part_ptrs, self._stage, self._stage_peers = [], [], [], []
        self._hdl = None
        for _ in range(self.N_BUF):
            t = symm_mem.empty(max(self.part_elems, 4), dtype=part_dtype, device=device)
            h = symm_mem.rendezvous(t, self.group)
            self._part.append(t)
            self._part_ptrs.append((ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs]))
            self._hdl = self._hdl or h
            s = symm_mem.empty(max(self.stage_elems, 4), dtype=stage_dtype, device=device)
            hs = symm_mem.rendezvous(s, self.group)
            self._stage.append(s)
            self._stage_peers.append([hs.get_buffer(p, (max(self.stage_elems, 4),), stage_dtype) for p in range(self.world)])
        self.stream = torch.cuda.Stream(device, priority=-1)
        self._n_part = 0          # partial buffers handed out so far
        self._n_stage = 0
        self._outstanding = None  # a partial buffer acquired and not yet reduced
        self._last_barrier = None # event: this rank is through the barrier of the latest peer collective

    # -- cross-rank ordering --------------------------------------------------------------------
    def _barrier(self):
        """On the communication stream: every rank has issued (and its earlier transfers have finished)."""
        self._hdl.barrier(channel=0)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self._last_barrier = ev

    # -- reduce-scatter + finish ---------------------------------------------------------------
    def acquire_partial(self, rows: int, feat: int, dtype):
        """A ``[rows, feat]`` view of the next symmetric partial buffer for the producer kernels of the
        compute stream (``out=``).  At most one buffer may be outstanding (acquired, not yet reduced)."""
        if self._outstanding is not None:
            raise _lib.TrgError("PeerComm: a partial buffer is already outstanding; reduce it first")
        if dtype != self.part_dtype or rows * feat > self.part_elems:
            raise _lib.TrgError(f"PeerComm: partial table [{rows}, {feat}] {dtype} does not fit the symmetric buffers "
                                f"({self.part_elems} x {self.part_dtype})")
        idx = self._n_part % self.N_BUF
        self._n_part += 1
        if self._last_barrier is not None:       # peers are done with this buffer once we are through that barrier
            torch.cuda.current_stream().wait_event(self._last_barrier)
        view = self._part[idx][:rows * feat].view(rows, feat)
        self._outstanding = (idx, view)
        return view

    def reduce_rows_async(self, part, out_dtype, row_scale=None, add=None, relu_of=None):
        """``finish(sum over ranks of part[owned rows])`` -> ``[rows / G, feat]`` of ``out_dtype``; ``part`` must
        be the outstanding buffer of ``acquire_partial``.  Runs on the communication stream after everything
        enqueued on the current stream so far (the producers, ``add``, ``relu_of``)."""
        if self._outstanding is None or part.data_ptr() != self._outstanding[1].data_ptr():
            raise _lib.TrgError("PeerComm.reduce_rows_async: not the outstanding partial buffer")
        idx, view = self._outstanding
        self._outstanding = None
        rows, feat = view.shape
        n = rows // self.world
        lib = _lib.load()
        out = torch.empty(n, feat, dtype=out_dtype, device=self.device)
        rs = row_scale.float().contiguous() if row_scale is not None else None
        for t, what in ((add, "add"), (relu_of, "relu_of")):
            if t is not None and (tuple(t.shape) != (n, feat) or t.dtype != out_dtype or not t.is_contiguous()):
                raise _lib.TrgError(f"PeerComm.reduce_rows_async: {what} must be a contiguous [{n}, {feat}] {out_dtype} tensor")
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            self._barrier()
            nbytes = n * feat * (self.world * view.element_size()
                                 + out.element_size() * (1 + (add is not None) + (relu_of is not None)))
            _lib.call("trg_peer_reduce_rows", nbytes, lib.trg_peer_reduce_rows, self._part_ptrs[idx], self.world,
                      self.rank * n, _lib.dtype_code(view.dtype), _lib.ptr(rs), _lib.ptr(add), _lib.ptr(relu_of), n, feat,
                      _lib.dtype_code(out_dtype), _lib.ptr(out), 0, self.stream.cuda_stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        for t in (out, rs, add, relu_of):
            if t is not None:
                t.record_stream(self.stream)
        return _Done(done, out, keep=(rs, add, relu_of))

    # -- all-gather ----------------------------------------------------------------------------------
    def all_gather_rows_async(self, x_local):
        """``[G * n, feat]`` = the owned rows of every rank, pulled by the copy engines."""
        x_local = x_local.contiguous()
        n, feat = x_local.shape
        ne = n * feat
        if x_local.dtype != self.stage_dtype or ne > self.stage_elems:
            raise _lib.TrgError(f"PeerComm: owned rows [{n}, {feat}] {x_local.dtype} do not fit the staging buffers")
        idx = self._n_stage % self.N_BUF
        self._n_stage += 1
        out = torch.empty(self.world * n, feat, dtype=x_local.dtype, device=self.device)
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            # staging copy: ordered after this rank's previous barrier on the same stream, i.e. after every
            # peer's pulls of the collective that used this staging buffer before
            self._stage[idx][:ne].copy_(x_local.view(-1))
            self._barrier()
            flat = out.view(-1)
            for k in range(self.world):
                p = (self.rank + k) % self.world           # start with the local chunk, then ring order
                flat[p * ne:(p + 1) * ne].copy_(self._stage_peers[idx][p][:ne])
            done = torch.cuda.Event()
            done.record(self.stream)
        out.record_stream(self.stream)
        x_local.record_stream(self.stream)
        return _Done(done, out, keep=x_local)


def peer_comm_for(shard, feat_max: int, table_dtype, xfer_dtype=None):
    """The ``PeerComm`` of a ``ShardedGraph`` (cached on it): buffers sized for ``[U_pad, feat_max]`` partial
    tables in the transport dtype and ``[users_per_rank, feat_max]`` staged rows."""
    key = (feat_max, table_dtype, xfer_dtype)
    pc = getattr(shard, "_peer_comm", None)
    if pc is None or pc[0] != key:
        dev = shard.x_local["user"].device
        pdt = xfer_dtype or table_dtype
        # staged rows: owned user rows (storage dtype) and their fp32... gradients keep the storage dtype too
        comm = PeerComm(shard.cu * shard.world * feat_max, pdt, shard.cu * feat_max, table_dtype, dev)
        shard._peer_comm = pc = (key, comm)
    return pc[1]
