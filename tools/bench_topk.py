"""Top-k recs/s (BASELINE config 5 shape): python tools/bench_topk.py [P] [H] [K] [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import truth_recommendation_gnn_b200 as trg
from truth_recommendation_gnn_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
dev = torch.device("cuda")
q, cat = synth.synth_queries(B, P, H, device=dev, dtype=torch.bfloat16)
trg.score_topk(q, cat, K); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 2
e0.record()
for _ in range(it): v, i = trg.score_topk(q, cat, K)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / it
print(f"B={B} P={P} H={H} K={K} bf16: {ms:.2f} ms/batch  {B/ms*1e3:,.0f} users/s  {B*K/ms*1e3:,.0f} recs/s  "
      f"{2.0*B*P*H/ms/1e9:.1f} TFLOP/s", flush=True)
# spot check against torch on a slice of queries
ref = (q[:8].float() @ cat.float().t()) if P <= 5_000_000 else None
if ref is not None:
    tv, ti = torch.topk(ref, K)
    print("values match torch.topk:", torch.allclose(tv, v[:8], rtol=1e-5, atol=1e-5))
