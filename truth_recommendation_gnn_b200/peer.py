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

    def __init__(self, part_elems: int, part_dtype, stage_elems: int, stage_dtype, device, group=None,
                 reduce_mode: str = "staged"):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device
        self.part_dtype, self.stage_dtype = part_dtype, stage_dtype
        self.part_elems, self.stage_elems = int(part_elems), int(stage_elems)
        self._part, self._part_ptrs, self._part_peers, self._stage, self._stage_peers = [], [], [], [], []
        self._hdl = None
        # "staged": the copy engines pull the peers' chunks of the owned row range into local scratch and the
        # reduce kernel reads local memory only -- no NVLink load ever waits in an SM next to the aggregation
        # kernel running meanwhile (measured at 8 GPUs: plain peer loads from 296 thin CTAs cost the concurrent
        # gathers 1 ms per step, NCCL's kernels 0.6 ms).  "direct": the kernel loads straight from the peers.
        if reduce_mode not in ("staged", "direct"):
            raise _lib.TrgError(f"PeerComm: unknown reduce_mode {reduce_mode!r}")
        self.reduce_mode = reduce_mode
        self._scratch = None
        for _ in range(self.N_BUF):
            t = symm_mem.empty(max(self.part_elems, 4), dtype=part_dtype, device=device)
            h = symm_mem.rendezvous(t, self.group)
            self._part.append(t)
            self._part_ptrs.append((ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs]))
            self._part_peers.append([h.get_buffer(p, (max(self.part_elems, 4),), part_dtype) for p in range(self.world)])
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

    def begin_step(self):
        """Start of a step (eager or inside a CUDA-graph capture): one barrier that every rank enters after
        ALL its transfers of the previous step, then the buffer rotation restarts from buffer 0.  A step --
        or a replay of its captured graph -- is thereby self-contained: which buffer a collective uses does not
        depend on the steps before it."""
        if self._outstanding is not None:
            raise _lib.TrgError("PeerComm.begin_step: a partial buffer is still outstanding")
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            self._barrier()
        self._n_part = self._n_stage = 0

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
            es = view.element_size()
            nbytes = n * feat * (self.world * es + out.element_size() * (1 + (add is not None) + (relu_of is not None)))
            if self.reduce_mode == "direct":
                ptrs, row0 = self._part_ptrs[idx], self.rank * n
            else:
                if self._scratch is None:      # [G - 1 chunks of the largest owned row range]
                    self._scratch = torch.empty((self.world - 1) * (self.part_elems // self.world), dtype=self.part_dtype,
                                                device=self.device)
                ne, lo = n * feat, self.rank * n * feat
                slots = [0] * self.world
                slots[self.rank] = self._part[idx].data_ptr() + lo * es
                for k in range(1, self.world):
                    p = (self.rank + k) % self.world          # ring order: every peer serves a different reader
                    dst = self._scratch[(k - 1) * ne:k * ne]
                    dst.copy_(self._part_peers[idx][p][lo:lo + ne])
                    slots[p] = dst.data_ptr()
                ptrs, row0 = (ctypes.c_void_p * self.world)(*slots), 0
            _lib.call("trg_peer_reduce_rows", nbytes, lib.trg_peer_reduce_rows, ptrs, self.world,
                      row0, _lib.dtype_code(view.dtype), _lib.ptr(rs), _lib.ptr(add), _lib.ptr(relu_of), n, feat,
                      _lib.dtype_code(out_dtype), _lib.ptr(out), 0 if self.reduce_mode == "direct" else 8,
                      self.stream.cuda_stream)
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


class HostSliceGather:
    """A host array that every rank holds (the step's sampled negatives, train_gnn.py:272) brought to every
    device with 1/G of the PCIe traffic: each rank uploads only its slice (pinned host -> symmetric buffer) and
    pulls the other slices from its peers over NVLink with the copy engines.  Eight ranks uploading the full
    320 MB array of config 2 at once made the end-to-end step host-bandwidth-bound (21.8 ms vs 13.9 ms
    resident); 40 MB per rank is 0.7 ms.  Own stream and own barrier channel: independent of the order of the
    step's other collectives.  Two rotating staging buffers, one barrier per call (see ``PeerComm``)."""

    def __init__(self, n_elems: int, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n, self.dtype, self.device = int(n_elems), dtype, device
        self.chunk = (self.n + self.world - 1) // self.world
        self._stage, self._peers, self._hdl = [], [], None
        for _ in range(2):
            t = symm_mem.empty(max(self.chunk, 4), dtype=dtype, device=device)
            h = symm_mem.rendezvous(t, self.group)
            self._stage.append(t)
            self._peers.append([h.get_buffer(p, (max(self.chunk, 4),), dtype) for p in range(self.world)])
            self._hdl = self._hdl or h
        self.stream = torch.cuda.Stream(device)
        self._calls = 0

    def gather_async(self, host, out=None, event=None):
        """-> (device tensor [n], event).  ``host``: the full array, identical on every rank; ``out``: a device
        tensor to fill (the static input of a captured step) instead of a fresh one; ``event``: the event to
        record when the array is complete (a captured step waits on an ``external`` event) instead of a new one."""
        if host.is_cuda or host.dtype != self.dtype or host.numel() != self.n:
            raise _lib.TrgError(f"HostSliceGather: expected a host {self.dtype} tensor of {self.n} elements")
        idx = self._calls % 2
        self._calls += 1
        if out is None:
            out = torch.empty(self.n, dtype=self.dtype, device=self.device)
        else:
            self.stream.wait_stream(torch.cuda.current_stream())      # earlier readers of `out` are done
        c = self.chunk
        lo, hi = min(self.rank * c, self.n), min((self.rank + 1) * c, self.n)
        with torch.cuda.stream(self.stream):
            if hi > lo:
                self._stage[idx][:hi - lo].copy_(host.view(-1)[lo:hi], non_blocking=True)
            self._hdl.barrier(channel=1)
            for k in range(self.world):
                p = (self.rank + k) % self.world
                a, b = min(p * c, self.n), min((p + 1) * c, self.n)
                if b > a:
                    out[a:b].copy_(self._peers[idx][p][:b - a])
            ev = event if event is not None else torch.cuda.Event()
            ev.record(self.stream)
        out.record_stream(self.stream)
        return out, ev


def peer_comm_for(shard, feat_max: int, table_dtype, xfer_dtype=None, reduce_mode="staged"):
    """The ``PeerComm`` of a ``ShardedGraph`` (cached on it): buffers sized for ``[U_pad, feat_max]`` partial
    tables in the transport dtype and ``[users_per_rank, feat_max]`` staged rows."""
    key = (feat_max, table_dtype, xfer_dtype, reduce_mode)
    pc = getattr(shard, "_peer_comm", None)
    if pc is None or pc[0] != key:
        dev = shard.x_local["user"].device
        pdt = xfer_dtype or table_dtype
        # staged rows: owned user rows (storage dtype) and their fp32... gradients keep the storage dtype too
        comm = PeerComm(shard.cu * shard.world * feat_max, pdt, shard.cu * feat_max, table_dtype, dev,
                        reduce_mode=reduce_mode)
        shard._peer_comm = pc = (key, comm)
    return pc[1]
