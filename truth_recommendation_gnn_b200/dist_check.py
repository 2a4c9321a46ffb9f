"""Parity of the destination-partitioned multi-GPU path against the single-GPU path, run INSIDE a
``torch.distributed`` job (every rank calls it): ``bench.py --gpus N`` reports it on its JSON line at every
N > 1 and ``tools/check_dist_gpu.py`` asserts it.  Both sides are the product's CUDA kernels; what is checked
is the partitioning, the collectives and their fp32 transport -- the single-GPU path itself is checked against
the CPU oracle by tests/.

Compared on identical weights (nothing here goes through an optimizer step first: Adam turns a noise-level
difference in a near-zero gradient into a full +-lr weight move, which would measure Adam, not the path):
embeddings of the owned rows, the loss, every weight gradient, and the sharded catalogue top-k (ids bit-equal).
Tolerances are BASELINE.json's: 1e-5 relative for fp32, 1e-2 for bf16 (relative to the tensor's scale:
reduce-scatter changes the summation order of signed sums)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import dist as tdist
from . import dist_fused, synth
from .functional import score_topk
from .nn import StackedWeightedRGCN
from .train import train_step

TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2}


def _err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def check_sharded_against_single(device, dtype=torch.float32, fused=True, sizes=(20_011, 70_003, 600_000, 150_000),
                                 hidden=128, layers=2, from_generator=False, graph=True):
    """Returns ``dict(emb_err, loss_err, grad_err, topk_ids_equal, ok, ...)``, identical on every rank."""
    U, P, EE, ES = sizes
    world = dist.get_world_size()
    if from_generator:
        cg = synth.CounterGraph(U, P, EE, ES, hidden, seed=0)
        g = cg.materialize(device, dtype)
        shard = tdist.ShardedGraph.from_generator(cg, device, dtype, chunk=200_000)
    else:
        g = synth.synth_graph(U, P, EE, ES, hidden, seed=0, device=device, dtype=dtype)
        shard = tdist.ShardedGraph(g.x_dict, g.edge_index_dict, g.train_edge_index, g.interaction_type_tensor, U, P)
    sd = synth.init_state_dict(hidden, hidden, layers)
    ref = StackedWeightedRGCN(hidden, layers)
    ref.load_state_dict(sd)
    ref = ref.to(device).to(dtype)
    mod = StackedWeightedRGCN(hidden, layers)
    mod.load_state_dict(sd)
    mod = mod.to(device).to(dtype)
    # forward on identical weights
    with torch.no_grad():
        full = ref(g.x_dict, g.edge_index_dict)
        loc = tdist.forward_sharded(mod, shard)
    nu, np_ = shard.u1 - shard.u0, shard.p1 - shard.p0
    emb_err = max(_err(loc["user"][:nu].float(), full["user"][shard.u0:shard.u1].float()) if nu > 0 else 0.0,
                  _err(loc["post"][:np_].float(), full["post"][shard.p0:shard.p1].float()) if np_ > 0 else 0.0)
    # one step from identical weights: loss + gradients (lr = 0 keeps the weights where they are)
    o_ref = torch.optim.SGD(ref.parameters(), lr=0.0)
    o_mod = torch.optim.SGD(mod.parameters(), lr=0.0)
    neg = synth.synth_neg(P, EE, 0, device=device)
    l_ref = train_step(ref, o_ref, g.x_dict, g.edge_index_dict, g.train_edge_index, g.interaction_type_tensor,
                       U, P, neg_p=neg)
    if fused:
        l_mod = dist_fused.train_step_sharded_fused(mod, o_mod, shard, neg_p_global=neg)
    else:
        l_mod = tdist.train_step_sharded(mod, o_mod, shard, neg_p_global=neg)
    loss_err = abs(l_mod - l_ref) / abs(l_ref)
    grad_err = max(_err(a.grad.float(), b.grad.float()) for a, b in zip(mod.parameters(), ref.parameters()))
    graph_err = None
    if fused and graph:
        # the same step replayed from its CUDA graph (first call captures, second replays): same loss, same grads
        o_g = torch.optim.SGD(mod.parameters(), lr=0.0)
        for _ in range(2):
            l_g = dist_fused.train_step_sharded_fused(mod, o_g, shard, neg_p_global=neg, cuda_graph=True)
        graph_err = max(abs(l_g - l_ref) / abs(l_ref),
                        max(_err(a.grad.float(), b.grad.float()) / 5 for a, b in zip(mod.parameters(), ref.parameters())))
        shard._graphed.close()
        shard._graphed = None
    # sharded catalogue top-k == unsharded (ids bit-equal: the same scores row by row)
    q = full["user"][:257].contiguous()
    ev, ei = score_topk(q, full["post"].contiguous(), 100)
    sv, si = tdist.recommend_sharded(q, full["post"][shard.p0:shard.p1].contiguous(), 100, shard.p0)
    ids_equal = bool(torch.equal(si, ei)) and bool(torch.equal(sv, ev))
    t = torch.tensor([emb_err, loss_err, grad_err, 0.0 if ids_equal else 1.0, graph_err or 0.0], device=device,
                     dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)              # worst rank
    emb_err, loss_err, grad_err, bad_ids, graph_err_w = (float(x) for x in t)
    tol = TOL[dtype]
    return dict(world=world, dtype=str(dtype).replace("torch.", ""), path="fused" if fused else "tape",
                graph=f"{U} users / {P} posts / {EE + ES} edges, H={hidden}, L={layers}"
                      + (" (counter-based, sharded without materialising)" if from_generator else ""),
                emb_err=emb_err, loss_err=loss_err, grad_err=grad_err, topk_ids_equal=bad_ids == 0.0, tol=tol,
                grad_tol=5 * tol,     # gradients: sums over all nodes of signed terms, as in tests/ (5 x tol)
                graph_replay_err=graph_err_w if graph_err is not None else None,    # max(loss err, grad err / 5)
                ok=bool(emb_err <= tol and loss_err <= tol and grad_err <= 5 * tol and bad_ids == 0.0
                        and graph_err_w <= tol), loss=l_mod)
