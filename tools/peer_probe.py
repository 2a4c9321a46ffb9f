"""Probe of the peer-memory plumbing on the box (torchrun, N ranks): symmetric-memory rendezvous, copy-engine
pull all-gather vs NCCL all-gather, torch-level peer-read sum vs NCCL reduce-scatter.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py [rows] [feat]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
feat = int(sys.argv[2]) if len(sys.argv) > 2 else 128
chunk = (rows + world - 1) // world


def log(*a):
    if rank == 0:
        print(*a, flush=True)


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / it], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


import torch.distributed._symmetric_memory as symm_mem

try:
    t0 = time.time()
    src = symm_mem.empty(chunk * feat, dtype=torch.float32, device=dev)
    h_src = symm_mem.rendezvous(src, dist.group.WORLD)
    part = symm_mem.empty(chunk * world * feat, dtype=torch.float32, device=dev)
    h_part = symm_mem.rendezvous(part, dist.group.WORLD)
    log(f"symm_mem ok in {time.time() - t0:.2f}s: world={h_src.world_size} multicast={h_src.has_multicast_support} "
        f"mc_ptr={h_part.multicast_ptr:#x} buffer_ptrs={[hex(p) for p in h_src.buffer_ptrs][:3]}")
except Exception as e:  # noqa: BLE001
    log("symm_mem FAILED:", repr(e))
    raise

src.copy_(torch.randn(chunk * feat, device=dev))
part.copy_(torch.randn(chunk * world * feat, device=dev))
full = torch.empty(world * chunk * feat, dtype=torch.float32, device=dev)
mb = chunk * feat * 4 / 1e6

# NCCL baselines
ms = timeit(lambda: dist.all_gather_into_tensor(full, src))
log(f"NCCL all_gather      {mb:.0f} MB/rank: {ms:.3f} ms  ingress {(world - 1) * mb / ms:.0f} GB/s")
out = torch.empty(chunk * feat, dtype=torch.float32, device=dev)
ms = timeit(lambda: dist.reduce_scatter_tensor(out, part))
log(f"NCCL reduce_scatter  {mb:.0f} MB/rank: {ms:.3f} ms  ingress {(world - 1) * mb / ms:.0f} GB/s")
ref_full = full.clone()
ref_out = out.clone()

# copy-engine pull all-gather: barrier, then one cudaMemcpyAsync per peer (1 stream / one stream per peer)
peers = [h_src.get_buffer(p, (chunk * feat,), torch.float32) for p in range(world)]
streams = [torch.cuda.Stream(dev) for _ in range(world)]


def ce_pull(nstreams):
    h_src.barrier(channel=0)
    if nstreams == 1:
        for k in range(world):
            p = (rank + k) % world
            full[p * chunk * feat:(p + 1) * chunk * feat].copy_(peers[p])
    else:
        ev = torch.cuda.Event(); ev.record()
        for k in range(world):
            p = (rank + k) % world
            s = streams[k % nstreams]
            s.wait_event(ev)
            with torch.cuda.stream(s):
                full[p * chunk * feat:(p + 1) * chunk * feat].copy_(peers[p])
        for s in streams[:nstreams]:
            e = torch.cuda.Event(); e.record(s); torch.cuda.current_stream().wait_event(e)


for ns in (1, 2, 4):
    if ns > world:
        continue
    full.zero_()
    ms = timeit(lambda: ce_pull(ns))
    ok = torch.equal(full, ref_full)
    log(f"CE pull all_gather ({ns} streams): {ms:.3f} ms  ingress {(world - 1) * mb / ms:.0f} GB/s  equal={ok}")

# peer-read sum through torch ops (what a fused kernel would do with plain loads)
pparts = [h_part.get_buffer(p, (world * chunk * feat,), torch.float32) for p in range(world)]


def peer_sum():
    h_part.barrier(channel=1)
    lo, hi = rank * chunk * feat, (rank + 1) * chunk * feat
    acc = pparts[0][lo:hi].clone()
    for p in range(1, world):
        acc += pparts[p][lo:hi]
    return acc


ms = timeit(peer_sum)
got = peer_sum()
log(f"torch peer-read sum: {ms:.3f} ms  ingress {(world - 1) * mb / ms:.0f} GB/s  max|diff| vs NCCL {(got - ref_out).abs().max().item():.3e}")
ms = timeit(lambda: h_src.barrier(channel=0), it=50)
log(f"symm_mem barrier: {ms * 1e3:.1f} us")
dist.barrier()
dist.destroy_process_group()
