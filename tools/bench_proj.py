"""Micro-benchmark of K3 (projection kernels) against cuBLAS via torch: python tools/bench_proj.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import functional as Fn, _lib

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

dev = torch.device("cuda")
for dtype in (torch.float32, torch.bfloat16):
    for n, h, nt in ((5_000_000, 128, 2), (1_000_000, 128, 3), (2_000_000, 256, 2), (4_000_000, 64, 2)):
        A = [torch.randn(n, h, device=dev).to(dtype) for _ in range(nt)]
        W = [(torch.randn(h, h, device=dev) / h ** 0.5).to(dtype) for _ in range(nt)]
        b = torch.randn(h, device=dev)
        es = A[0].element_size()
        bytes_ = n * (nt * h + h) * es
        flops = 2.0 * n * nt * h * h
        t = timeit(lambda: Fn.sage_proj_fwd([(a, w, 1.0) for a, w in zip(A, W)], b, True))
        def ref():
            o = torch.addmm(b.to(dtype), A[0], W[0].t())
            for a, w in zip(A[1:], W[1:]): o.addmm_(a, w.t())
            return torch.relu_(o)
        tr = timeit(ref)
        dz = torch.randn(n, h, device=dev).to(dtype)
        tb = timeit(lambda: Fn.sage_proj_bwd_input(dz, [(w, 1.0, None) for w in W]))
        print(f"{str(dtype):15s} n={n:8d} h={h:3d} terms={nt}: fwd {t:7.3f} ms  {bytes_/t/1e6:7.0f} GB/s  {flops/t/1e9:6.1f} TF/s"
              f" | torch {tr:7.3f} ms | bwd_input {tb:7.3f} ms {bytes_/tb/1e6:7.0f} GB/s", flush=True)
        del A, dz
