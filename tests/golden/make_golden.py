"""Generates the committed golden fixtures from the CPU oracle (run in the build container):

    python tests/golden/make_golden.py

The reference ships no golden vectors and its operator library (torch_geometric) is not
installable here, so these vectors pin the ORACLE's outputs (oracle/__init__.py: "parity
unpinned" for SAGEConv); the torch-only parts (mm/topk values, BCE, Adam) are executed by real
torch.  Fixtures are small .pt files holding inputs AND expected outputs.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import csr as ocsr  # noqa: E402
from oracle import sage as osage  # noqa: E402
from oracle import topk as otopk  # noqa: E402
from truth_recommendation_gnn_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(name, u, p, e_eng, e_soc, h, layers, steps, k, skew=False, extra_edges=None):
    torch.set_num_threads(1)  # fixed summation order
    g = synth.synth_graph(u, p, e_eng, e_soc, h, seed=0, skew=skew)
    if extra_edges is not None:
        g = extra_edges(g)
    sd = synth.init_state_dict(h, h, layers, seed=1)
    model = (osage.WeightedRGCNOracle(h, (h, h)) if layers == 1
             else osage.StackedWeightedRGCNOracle(h, layers, (h, h)))
    model.load_state_dict(sd)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    with torch.no_grad():
        out0 = model(g.x_dict, g.edge_index_dict)
    losses, negs = [], []
    for s in range(steps):
        neg = synth.synth_neg(p, g.train_edge_index.size(1), s)
        negs.append(neg)
        losses.append(osage.train_step(model, opt, g.x_dict, g.edge_index_dict, g.train_edge_index,
                                       g.interaction_type_tensor, u, p, neg_p=neg))
    grads = {n: prm.grad.clone() for n, prm in model.named_parameters()}
    with torch.no_grad():
        out1 = model(g.x_dict, g.edge_index_dict)
    vals, ids = otopk.score_topk(out1["user"], out1["post"], k)
    csr = {str(rel): ocsr.csr_by_dst(ei, g.x_dict[rel[2]].size(0)) for rel, ei in g.edge_index_dict.items()}
    fix = dict(
        meta=dict(u=u, p=p, h=h, layers=layers, steps=steps, k=k),
        x_user=g.x_dict["user"], x_post=g.x_dict["post"],
        edge_index={str(k_): v for k_, v in g.edge_index_dict.items()},
        w=g.interaction_type_tensor, state_dict=sd, neg=negs,
        out0_user=out0["user"], out0_post=out0["post"],
        losses=torch.tensor(losses, dtype=torch.float64),
        last_grads=grads, state_dict_after={k_: v.detach().clone() for k_, v in model.state_dict().items()},
        out1_user=out1["user"], out1_post=out1["post"], topk_vals=vals, topk_ids=ids,
        csr={r: dict(rowptr=c[0].int(), col=c[1].int(), eid=c[2].int()) for r, c in csr.items()},
    )
    path = os.path.join(HERE, name + ".pt")
    torch.save(fix, path)
    print(name, os.path.getsize(path) // 1024, "KiB", "losses", losses)


def add_edge_cases(g):
    """duplicates, a self loop in social, an isolated user/post (ids U-1 / P-1 never a destination)."""
    eng = g.edge_index_dict[synth.REL_ENGAGE]
    soc = g.edge_index_dict[synth.REL_SOCIAL]
    u, p = g.num_users, g.num_posts
    eng = eng[:, (eng[0] != u - 1) & (eng[1] != p - 1)]
    soc = soc[:, (soc[1] != u - 1)]
    eng = torch.cat([eng, eng[:, :3], eng[:, :1]], dim=1)          # duplicate edges (kept, not coalesced)
    soc = torch.cat([soc, torch.tensor([[2, 0], [2, 0]])], dim=1)  # self loop + extra
    ei = {synth.REL_SOCIAL: soc.contiguous(), synth.REL_ENGAGE: eng.contiguous(),
          synth.REL_DIRECT: eng.flip(0).contiguous()}
    return synth.SynthGraph(g.x_dict, ei, ei[synth.REL_ENGAGE], g.interaction_type_tensor, u, p)


if __name__ == "__main__":
    run_case("tiny_l1", 7, 9, 30, 12, 8, 1, 2, 3, extra_edges=add_edge_cases)
    run_case("small_l2", 150, 400, 3000, 800, 16, 2, 3, 10)
    run_case("small_l1_skew", 120, 300, 2500, 600, 64, 1, 2, 10, skew=True)
