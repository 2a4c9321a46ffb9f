"""Load-balance report: aggregation kernels on the uniform vs the Zipf-like (dst = floor(N*u^3)) graph."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import truth_recommendation_gnn_b200 as trg
from truth_recommendation_gnn_b200 import functional as Fn, synth
dev = torch.device("cuda")
U, P, EE, ES, H = 1_000_000, 5_000_000, 40_000_000, 10_000_000, 128
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for skew in (False, True):
    g = synth.synth_graph(U, P, EE, ES, H, device=dev, skew=skew)
    for rel in (synth.REL_DIRECT, synth.REL_SOCIAL, synth.REL_ENGAGE):
        ei = g.edge_index_dict[rel]
        xs, nd = g.x_dict[rel[0]], g.x_dict[rel[2]].size(0)
        rg = trg.relation_graph(ei, xs.size(0), nd)
        csr = rg.fwd
        deg = (csr.rowptr[1:] - csr.rowptr[:-1])
        t = timeit(lambda: Fn.sage_agg_fwd(csr, xs, want_inv_deg=False))
        b = csr.n_edges * (H * 4 + 4) + 4 * (nd + 1) + nd * H * 4
        print(f"skew={skew} agg_fwd {rel[1]:12s} max_deg={int(deg.max()):8d} : {t:8.3f} ms {b/t/1e6:7.0f} GB/s", flush=True)
    del g; trg.clear_cache(); torch.cuda.empty_cache()
