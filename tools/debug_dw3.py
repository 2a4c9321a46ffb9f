import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import functional as Fn
dev = torch.device("cuda")
DT = torch.bfloat16 if "bf16" in sys.argv else torch.float32
h, n = 128, 64
for nterms in (0, 1):
    res = []
    for r0 in range(n):
        dz = torch.zeros(n, h); dz[r0, :] = 1.0
        terms = [(torch.full((n, 64), 0.5, device=dev, dtype=DT), 1.0) for _ in range(nterms)]
        outs, db = Fn.sage_proj_bwd_weight(dz.to(dev).to(DT), terms, True)
        res.append(float(db[0]))
    print("terms", nterms, "weight of each row in db:", res)
