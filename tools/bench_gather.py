"""Micro-benchmark of the gather kernels at config-2 shapes: python tools/bench_gather.py [bf16]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import truth_recommendation_gnn_b200 as trg
from truth_recommendation_gnn_b200 import functional as Fn, synth

dev = torch.device("cuda")
dtype = torch.bfloat16 if "bf16" in sys.argv else torch.float32
U, P, EE, ES, H = 1_000_000, 5_000_000, 40_000_000, 10_000_000, 128
g = synth.synth_graph(U, P, EE, ES, H, device=dev, dtype=dtype)
es = 2 if dtype == torch.bfloat16 else 4

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for rel in (synth.REL_DIRECT, synth.REL_SOCIAL, synth.REL_ENGAGE):
    ei = g.edge_index_dict[rel]
    xs, nd = g.x_dict[rel[0]], g.x_dict[rel[2]].size(0)
    rg = trg.relation_graph(ei, xs.size(0), nd)
    csr = rg.fwd
    t = timeit(lambda: Fn.sage_agg_fwd(csr, xs, want_inv_deg=False))
    b = csr.n_edges * (H * es + 4) + 4 * (nd + 1) + nd * H * es
    print(f"agg_fwd {rel[1]:12s} rows={nd:8d} E={csr.n_edges:9d}: {t:7.3f} ms {b/t/1e6:7.0f} GB/s", flush=True)
    gm = torch.randn(nd, H, device=dev).to(dtype)
    csr_t = rg.bwd
    inv = torch.rand(nd, device=dev)
    t = timeit(lambda: Fn.sage_agg_bwd(csr_t, inv, gm))
    ns = xs.size(0)
    b = csr.n_edges * (H * es + 4) + 4 * (ns + 1) + ns * H * es + 4 * nd
    print(f"agg_bwd {rel[1]:12s} rows={ns:8d}: {t:7.3f} ms {b/t/1e6:7.0f} GB/s   (no scale: {timeit(lambda: Fn.sage_agg_bwd(csr_t, None, gm)):7.3f} ms)", flush=True)
# loss-side kernels
u = torch.relu(torch.randn(U, H, device=dev)).to(dtype).requires_grad_(True)
p = torch.relu(torch.randn(P, H, device=dev)).to(dtype).requires_grad_(True)
neg = synth.synth_neg(P, EE, 0, device=dev)
ls = Fn.link_structure(g.train_edge_index, g.interaction_type_tensor, U, P)
t = timeit(lambda: Fn.edge_bce_fwd(ls, u.detach(), p.detach(), neg, True))
b = EE * (2 * H * es + 24) + U * (2 * H * es + 4)
print(f"edge_bce fwd+grad: {t:7.3f} ms {b/t/1e6:7.0f} GB/s", flush=True)
coef = torch.randn(EE, device=dev)
t = timeit(lambda: Fn.gather_wsum(ls.by_post, coef, u.detach()))
b = EE * (H * es + 12) + 4 * (P + 1) + P * H * es
print(f"wsum by_post: {t:7.3f} ms {b/t/1e6:7.0f} GB/s", flush=True)
t = timeit(lambda: trg.build_csr(g.train_edge_index[0], neg, P, U, validate=False))
print(f"neg csr build: {t:7.3f} ms", flush=True)
# single-row anchored loss launches (user-anchored, posts gathered): pos + neg as two passes
eid_long = ls.by_user.eid.long()
def two_pass():
    col_neg = neg.index_select(0, eid_long).int()
    csr_neg = trg.CSR(ls.by_user.rowptr, col_neg, ls.by_user.eid, ls.by_user.n_rows, ls.by_user.n_cols)
    l1, c1, gu = Fn.edge_anchor_loss(ls.by_user, u.detach(), p.detach(), EE, 1, ls.wbar, True, None)
    l2, c2, gu = Fn.edge_anchor_loss(csr_neg, u.detach(), p.detach(), EE, 0, ls.wbar, True, gu)
    return l1 + l2
t = timeit(two_pass)
print(f"edge loss as two single-row passes (incl. neg col gather): {t:7.3f} ms", flush=True)
t = timeit(lambda: Fn.edge_anchor_loss(ls.by_user, u.detach(), p.detach(), EE, 1, ls.wbar, True, None))
b = EE * (H * es + 12) + U * (2 * H * es + 4)
print(f"  one pass: {t:7.3f} ms {b/t/1e6:7.0f} GB/s", flush=True)
lf, _, _, _ = Fn.edge_bce_fwd(ls, u.detach(), p.detach(), neg, True)
print("  loss fused vs two-pass:", float(lf), float(two_pass()))
