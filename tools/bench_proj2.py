"""Projection kernels at cfg2 shapes (fwd / bwd_input / bwd_weight), fp32: python tools/bench_proj2.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import functional as Fn
dev = torch.device("cuda")
dtype = torch.bfloat16 if "bf16" in sys.argv else torch.float32
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
h = 128
es = 2 if dtype == torch.bfloat16 else 4
for n, nt, tag in ((5_000_000, 2, "posts"), (1_000_000, 3, "users")):
    A = [torch.randn(n, h, device=dev).to(dtype) for _ in range(nt)]
    W = [(torch.randn(h, h, device=dev) / h ** 0.5).to(dtype) for _ in range(nt)]
    b = torch.randn(h, device=dev)
    dz = torch.randn(n, h, device=dev).to(dtype)
    rs = torch.rand(n, device=dev)
    t = timeit(lambda: Fn.sage_proj_fwd([(a, w, 1.0) for a, w in zip(A, W)], b, True))
    by = n * (nt * h + h) * es
    print(f"{tag} fwd        : {t:7.3f} ms {by/t/1e6:6.0f} GB/s", flush=True)
    t = timeit(lambda: Fn.sage_proj_bwd_input(dz, [(w, 1.0, rs) for w in W]))
    print(f"{tag} bwd_input  : {t:7.3f} ms {by/t/1e6:6.0f} GB/s", flush=True)
    t = timeit(lambda: Fn.sage_proj_bwd_weight(dz, [(a, 1.0) for a in A], True))
    print(f"{tag} bwd_weight : {t:7.3f} ms {by/t/1e6:6.0f} GB/s", flush=True)
    del A, dz
