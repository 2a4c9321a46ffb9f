"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: destination partitioning, the
all-gather / reduce-scatter autograd pair, weight-gradient all-reduce, loss partitioning and the
sharded top-k merge.  Compute is injected from the oracle (tests/oracle_ops.py); the result must
equal the single-process oracle on the same graph."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sage as osage
from oracle import topk as otopk
from tests import oracle_ops
from tests.util import oracle_model
from truth_recommendation_gnn_b200 import dist as tdist
from truth_recommendation_gnn_b200 import synth
import truth_recommendation_gnn_b200 as trg

U, P, EE, ES, H, L = 101, 257, 2000, 500, 16, 2     # sizes that do NOT divide by the world size


def _worker(rank, world, port, q, fused=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        torch.manual_seed(1000 + rank)        # the ranks' default generators are unrelated, as in a real launch
        g = synth.synth_graph(U, P, EE, ES, H, seed=0)
        shard = tdist.ShardedGraph(g.x_dict, g.edge_index_dict, g.train_edge_index,
                                   g.interaction_type_tensor, U, P)
        model = trg.StackedWeightedRGCN(H, L)
        model.load_state_dict(synth.init_state_dict(H, H, L))
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        losses = []
        for s in range(2):
            neg = synth.synth_neg(P, EE, s)
            if fused:     # tape-free step with async (overlapped) collectives
                from truth_recommendation_gnn_b200 import dist_fused
                assert dist_fused.eligible(model, shard)
                losses.append(dist_fused.train_step_sharded_fused(model, opt, shard, neg_p_global=neg,
                                                                  prims=oracle_ops.OracleStepPrims))
            else:
                losses.append(tdist.train_step_sharded(model, opt, shard, neg_p_global=neg,
                                                       ops=oracle_ops.OracleOps, loss_ops=oracle_ops.OracleLossOps))
        with torch.no_grad():
            out = tdist.forward_sharded(model, shard, oracle_ops.OracleOps)
        # sharded recommendation over this rank's posts
        qv, cat = synth.synth_queries(7, P, H, zero_frac=0.2)
        cat_local = cat[shard.p0:shard.p1].contiguous()
        rv, ri = tdist.recommend_sharded(qv, cat_local, 10, shard.p0, oracle_ops.score_topk, oracle_ops.merge)
        # a counter-based graph sharded WITHOUT materialising it == the partition of the materialised graph
        cg = synth.CounterGraph(U, P, EE, ES, H, seed=3)
        gm = cg.materialize()
        sa = tdist.ShardedGraph(gm.x_dict, gm.edge_index_dict, gm.train_edge_index, gm.interaction_type_tensor, U, P)
        sb = tdist.ShardedGraph.from_generator(cg, "cpu", chunk=777)
        for t in ("user", "post"):
            assert torch.equal(sa.x_local[t], sb.x_local[t]), t
        for rel in sa.rels:
            ra, rb = sa.rels[rel], sb.rels[rel]
            if isinstance(ra, trg.graph.PushRelation):
                assert torch.equal(ra.inv_deg, rb.inv_deg)
                ra, rb = ra.rel, rb.rel
            assert torch.equal(ra.edge_index, rb.edge_index) and (ra.n_src, ra.n_dst) == (rb.n_src, rb.n_dst), rel
        assert torch.equal(sa.pos_local, sb.pos_local) and torch.equal(sa.pos_u_global, sb.pos_u_global)
        assert torch.equal(sa.wbar, sb.wbar) and sa.n_pos_global == sb.n_pos_global
        # negatives drawn inside the step (neg_p_global=None) must be ONE array shared by all ranks
        n1, n2 = shard.draw_negatives(), shard.draw_negatives()
        assert n1.shape == (EE,) and int(n1.min()) >= 0 and int(n1.max()) < P and not torch.equal(n1, n2)
        # numpy payloads: tensors sent through mp queues need the sender alive until they are read
        q.put((rank, losses, {k: v.detach().numpy().copy() for k, v in model.state_dict().items()},
               out["user"][:shard.u1 - shard.u0].numpy().copy(), out["post"][:shard.p1 - shard.p0].numpy().copy(),
               rv.numpy().copy(), ri.numpy().copy(), torch.stack([n1, n2]).numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fused", [False, True])
def test_sharded_train_step_and_topk_match_single_process(fused):
    world, port = 2, 29500 + os.getpid() % 2000 + (7 if fused else 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, fused)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    t = torch.from_numpy
    negs = [t(r[7]) for r in res]
    assert all(torch.equal(n, negs[0]) for n in negs), "ranks drew different negatives"
    res = [(r[0], r[1], {k: t(v) for k, v in r[2].items()}, t(r[3]), t(r[4]), t(r[5]), t(r[6])) for r in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    torch.set_num_threads(1)
    g = synth.synth_graph(U, P, EE, ES, H, seed=0)
    ref = oracle_model(H, L, synth.init_state_dict(H, H, L))
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    ref_losses = [osage.train_step(ref, opt, g.x_dict, g.edge_index_dict, g.train_edge_index,
                                   g.interaction_type_tensor, U, P, neg_p=synth.synth_neg(P, EE, s)) for s in range(2)]
    with torch.no_grad():
        ref_out = ref(g.x_dict, g.edge_index_dict)
    for rank, losses, sd, u_loc, p_loc, rv, ri in res:
        for a, b in zip(losses, ref_losses):
            assert abs(a - b) <= 1e-5 * abs(b), (rank, a, b)          # global loss on every rank
        for k, v in ref.state_dict().items():
            assert torch.allclose(sd[k], v, atol=2e-6), k             # identical weights after 2 steps
    u = torch.cat([r[3] for r in res])
    p = torch.cat([r[4] for r in res])
    assert u.shape == ref_out["user"].shape and p.shape == ref_out["post"].shape
    assert torch.allclose(u, ref_out["user"], atol=1e-5) and torch.allclose(p, ref_out["post"], atol=1e-5)
    qv, cat = synth.synth_queries(7, P, H, zero_frac=0.2)
    ev, ei = otopk.score_topk(qv, cat, 10)
    for r in res:
        assert torch.equal(r[6], ei) and torch.equal(r[5], ev)        # sharded top-k == unsharded, bit-exact ids


def test_partition_covers_every_edge_once():
    g = synth.synth_graph(U, P, EE, ES, H, seed=0)
    for world in (2, 3, 8):
        tot = {rel: 0 for rel in g.edge_index_dict}
        pos = 0
        for r in range(world):
            s = tdist.ShardedGraph.__new__(tdist.ShardedGraph)
            s.rank, s.world = r, world
            cu, cp = tdist.chunk_of(U, world), tdist.chunk_of(P, world)
            for rel, ei in g.edge_index_dict.items():
                c, n = (cu, U) if rel[2] == "user" else (cp, P)
                a, b = r * c, min((r + 1) * c, n)
                tot[rel] += int(((ei[1] >= a) & (ei[1] < b)).sum())
            a, b = r * cu, min((r + 1) * cu, U)
            pos += int(((g.train_edge_index[0] >= a) & (g.train_edge_index[0] < b)).sum())
        assert all(tot[rel] == g.edge_index_dict[rel].size(1) for rel in tot) and pos == EE
