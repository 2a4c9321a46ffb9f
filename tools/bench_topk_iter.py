"""Per-iteration device time of the config-5 top-k call with a host sync between calls (what the e2e leg does)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import truth_recommendation_gnn_b200 as trg
from truth_recommendation_gnn_b200 import synth
dev = torch.device("cuda")
q, cat = synth.synth_queries(4096, 50_000_000, 128, device=dev, dtype=torch.bfloat16)
qh = q.cpu().pin_memory()
trg.score_topk(q, cat, 100); torch.cuda.synchronize()
for mode in ("sync-between", "back-to-back", "e2e"):
    ts = []
    t0 = time.perf_counter()
    evs = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "e2e":
            v, ids = trg.score_topk(qh.to(dev, non_blocking=True), cat, 100)
            e1.record()
            v, ids = v.cpu(), ids.cpu()
        else:
            v, ids = trg.score_topk(q, cat, 100)
            e1.record()
            if mode == "sync-between": torch.cuda.synchronize()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / 4
    print(mode, [round(a.elapsed_time(b), 1) for a, b in evs], f"wall/iter {wall:.1f} ms", flush=True)
