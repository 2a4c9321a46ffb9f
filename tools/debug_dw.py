import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import functional as Fn
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
dev = torch.device("cuda")
DT = torch.bfloat16 if "bf16" in sys.argv else torch.float32
n, h = 32, 128
for name, dz in (("col-id", (torch.arange(h).float() + 1)[None, :].repeat(n, 1)),
                 ("row-id", (torch.arange(n).float() + 1)[:, None].repeat(1, h)),
                 ("onehot r3", torch.zeros(n, h).index_fill_(0, torch.tensor([3]), 1.0) * (torch.arange(h).float() + 1)[None, :]),
                 ("onehot r11", torch.zeros(n, h).index_fill_(0, torch.tensor([11]), 1.0) * (torch.arange(h).float() + 1)[None, :])):
    _, db = Fn.sage_proj_bwd_weight(dz.to(dev).to(DT), [], True)
    print(name, "expected", dz.sum(0)[:8].tolist(), "...")
    print(db.cpu().reshape(4, 32))
for trial in range(3):
    dz = (torch.arange(h).float() + 1)[None, :].repeat(48, 1)
    a = torch.ones(48, 64)
    (dw,), db = Fn.sage_proj_bwd_weight(dz.to(dev).to(DT), [(a.to(dev).to(DT), 1.0)], True)
    print("trial", trial, "db[:6]", db.cpu()[:6].tolist(), "expected", dz.sum(0)[:6].tolist(), "dw[:3,0]", dw.cpu()[:3, 0].tolist())
# dW: A = one-hot columns
dz = torch.zeros(n, h); dz[:, 5] = 1.0            # only h=5 active, all rows
a = (torch.arange(128).float() + 1)[None, :].repeat(n, 1)   # a[r, f] = f+1
(dw,), _ = Fn.sage_proj_bwd_weight(dz.to(dev).to(DT), [(a.to(dev).to(DT), 1.0)], False)
print("dW row5 expected 16*(f+1):", dw.cpu()[5, :40])
print("nonzero rows:", dw.cpu().abs().sum(1).nonzero().flatten().tolist())
