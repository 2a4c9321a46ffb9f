import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import functional as Fn
dev = torch.device("cuda")
DT = torch.bfloat16 if "bf16" in sys.argv else torch.float32
h = 128
def run(n, nterms, tag):
    dz = (torch.arange(h).float() + 1)[None, :].repeat(n, 1)
    terms = [(torch.full((n, 64), 0.5, device=dev, dtype=DT), 1.0) for _ in range(nterms)]
    outs, db = Fn.sage_proj_bwd_weight(dz.to(dev).to(DT), terms, True)
    ok = torch.allclose(db.cpu(), dz.sum(0))
    print(tag, "n", n, "terms", nterms, "db ok" if ok else f"db WRONG {db.cpu()[:4].tolist()} exp {dz.sum(0)[:4].tolist()}", flush=True)
if "warm" in sys.argv:
    run(48, 1, "warm")
for i in range(3):
    run(32, 0, f"bias-only #{i}")
run(64, 0, "bias-only n64")
run(4096, 0, "bias-only n4096")
run(32, 2, "2 terms")
run(32, 0, "bias-only again")
