import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from truth_recommendation_gnn_b200 import functional as Fn
dev = torch.device("cuda")
n, h, nt = 128 * 400, 128, 2
A = [torch.zeros(n, 128, device=dev) for _ in range(nt)]
for t in range(nt):
    for j in range(4):
        A[t][:, 32 * j + (torch.arange(n, device=dev) % 32)] = 0  # placeholder
        A[t][torch.arange(n, device=dev), 32 * j + (torch.arange(n, device=dev) % 32)] = float(2 ** (t * 4 + j))
W = [torch.ones(h, 128, device=dev) for _ in range(nt)]
exp = 255.0
from collections import Counter
for rep in range(3):
    out = Fn.sage_proj_fwd([(a, w, 1.0) for a, w in zip(A, W)], None, False)
    bad = (out != exp).any(1).nonzero().flatten()
    vals = Counter((out[bad] - exp).flatten().tolist())
    rows_in_tile = Counter((bad % 128).tolist())
    cols = Counter(((out != exp).nonzero()[:, 1]).tolist())
    print(f"rep {rep}: bad rows {bad.numel()}, delta values {vals.most_common(8)}")
    print("   rows-in-tile histogram (row:count):", sorted(rows_in_tile.items())[:40])
    print("   distinct bad cols:", len(cols), "tiles:", sorted(set((bad // 128).tolist()))[:10])
    # within a bad row: are all columns equal?
    if bad.numel():
        r = out[bad[0]]
        print("   first bad row", int(bad[0]), "unique vals", torch.unique(r).tolist()[:8])
# larger soak: many tiles per CTA, 2 and 3 terms, repeated
torch.manual_seed(0)
for nt in (2, 3):
    n = 128 * 4000
    A = [torch.randn(n, 128, device=dev) for _ in range(nt)]
    W = [torch.randn(h, 128, device=dev) / 11 for _ in range(nt)]
    ref = sum(a.double() @ w.double().t() for a, w in zip(A, W))
    for rep in range(5):
        out = Fn.sage_proj_fwd([(a, w, 1.0) for a, w in zip(A, W)], None, False)
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        dz = Fn.sage_proj_bwd_input(A[0], [(w, 1.0, None) for w in W])
        e2 = max(float((d.double() - A[0].double() @ w.double()).abs().max() / ref.abs().max()) for d, w in zip(dz, W))
        print(f"soak terms={nt} rep={rep}: fwd err {err:.2e} bwd_input err {e2:.2e}", flush=True)
