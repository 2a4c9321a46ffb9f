"""At-scale parity: the configurations bench.py TIMES (BASELINE config 2: 1M users / 5M posts / 50M edges,
H = 128; config 5: 4096 queries x 50M posts, K = 100) checked at their full sizes.

The CPU oracle cannot run a whole config-2 step in test time, so it recomputes a random SAMPLE of the
outputs from the COO edge list (chunked oracle: the edges of ~10k sampled destination rows per relation,
100k sampled loss edges) and every sampled value is compared -- bit for bit for CSR slices and fp32
CSR-order means, 1e-5 for the layer output and the loss coefficients.  Whole-job scalars (the loss) are
compared with the reference's literal torch expression (train_gnn.py:259-281) evaluated chunk by chunk.
The catalogue sweep uses integer-valued bf16 inputs, for which every score is exact in fp32, so ids AND
values must equal an independent exact ranking of all 50M posts under (score desc, id asc)."""
import pytest
import torch
import torch.nn.functional as F

import truth_recommendation_gnn_b200 as trg
from oracle import csr as ocsr
from oracle import sage as osage
from tests.util import TOL_F32, assert_close, assert_close_elementwise
from truth_recommendation_gnn_b200 import functional as Fn
from truth_recommendation_gnn_b200 import synth

pytestmark = pytest.mark.gpu

CFG2 = dict(U=1_000_000, P=5_000_000, Ee=40_000_000, Es=10_000_000, H=128)


def _sub_coo(ei_cpu, rows):
    """Edges whose destination is in ``rows`` (sorted unique), in original order, with the destination
    relabelled to its position in ``rows``: the oracle then runs on a problem with len(rows) destinations."""
    pos = torch.searchsorted(rows, ei_cpu[1]).clamp(max=rows.numel() - 1)
    sel = rows[pos] == ei_cpu[1]
    eidx = sel.nonzero().flatten()
    return torch.stack([ei_cpu[0][eidx], pos[eidx]]), eidx


@pytest.mark.timeout(1500)
def test_config2_sampled_rows_and_full_loss(dev):
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    c = CFG2
    U, P, H = c["U"], c["P"], c["H"]
    g = synth.synth_graph(U, P, c["Ee"], c["Es"], H, seed=0, device=dev)
    sd = synth.init_state_dict(H, H, 1)
    model = trg.WeightedRGCN(H)
    model.load_state_dict(sd)
    model = model.to(dev)
    ref = osage.WeightedRGCNOracle(H, (H, H))
    ref.load_state_dict(sd)
    with torch.no_grad():
        out = model(g.x_dict, g.edge_index_dict)
    x_cpu = {k: v.cpu() for k, v in g.x_dict.items()}
    gen = torch.Generator().manual_seed(5)
    n_sample = 10_000
    sub_ei, rows_of = {}, {}
    for rel, ei in g.edge_index_dict.items():
        n_src, n_dst = g.x_dict[rel[0]].size(0), g.x_dict[rel[2]].size(0)
        ei_cpu = ei.cpu()
        rows = torch.unique(torch.randint(0, n_dst, (n_sample,), generator=gen))
        sub, eidx = _sub_coo(ei_cpu, rows)
        sub_ei[rel], rows_of[rel] = sub, rows
        # --- K0 at 40M / 10M edges: the CSR slices of the sampled rows, bit for bit (3 radix passes) ---
        rp, col, eid = ocsr.csr_by_dst(sub, rows.numel())
        rg = trg.relation_graph(ei, n_src, n_dst)
        grp = rg.fwd.rowptr.cpu().long()
        deg = grp[rows + 1] - grp[rows]
        assert torch.equal(deg, rp[1:] - rp[:-1]), rel
        assert int(grp[-1]) == ei.size(1) and bool((grp[1:] >= grp[:-1]).all())
        idx = torch.repeat_interleave(grp[rows], deg) + (torch.arange(int(deg.sum())) - torch.repeat_interleave(rp[:-1], deg))
        assert torch.equal(rg.fwd.col.cpu().long()[idx], col), rel
        assert torch.equal(rg.fwd.eid.cpu().long()[idx], eidx[eid]), rel          # original edge positions, stable
        # --- K1: fp32 CSR-order mean of the sampled rows, bit for bit ---
        mean_gpu, _ = Fn.sage_agg_fwd(rg.fwd, g.x_dict[rel[0]], want_inv_deg=False)
        mean_ref, _ = osage.scatter_mean(x_cpu[rel[0]].index_select(0, sub[0]), sub[1], rows.numel())
        assert torch.equal(mean_gpu[rows.to(dev)].cpu(), mean_ref), rel
        del mean_gpu, ei_cpu
    # --- layer output (K1 + K3 + combine + ReLU) of sampled users / posts against the oracle ---
    with torch.no_grad():
        ru = rows_of[synth.REL_DIRECT]
        # the two user relations must be evaluated on the SAME destination rows: re-sample social on ru
        soc_sub, _ = _sub_coo(g.edge_index_dict[synth.REL_SOCIAL].cpu(), ru)
        d, s, p = ref.msg_direct, ref.msg_social, ref.post_update
        xu, xp = x_cpu["user"], x_cpu["post"]
        exp_u = F.relu(1.0 * osage.sage_conv(xp, xu[ru], sub_ei[synth.REL_DIRECT], d.lin_l.weight, d.lin_l.bias, d.lin_r.weight)
                       + 0.75 * osage.sage_conv(xu, xu[ru], soc_sub, s.lin_l.weight, s.lin_l.bias, s.lin_r.weight))
        rp_ = rows_of[synth.REL_ENGAGE]
        exp_p = F.relu(osage.sage_conv(xu, xp[rp_], sub_ei[synth.REL_ENGAGE], p.lin_l.weight, p.lin_l.bias, p.lin_r.weight))
    assert_close(out["user"][ru.to(dev)].cpu(), exp_u, TOL_F32, "config-2 user rows")
    assert_close(out["post"][rp_.to(dev)].cpu(), exp_p, TOL_F32, "config-2 post rows")
    # --- loss: full-size kernel result vs the literal torch expression, chunk by chunk ---
    neg = synth.synth_neg(P, c["Ee"], 0, device=dev)
    ue, pe = out["user"], out["post"]
    pos_u, pos_p = g.train_edge_index
    ls = Fn.link_structure(g.train_edge_index, g.interaction_type_tensor, U, P)
    # the product path (fused_step / LinkBCEFn): two single-row anchored passes over the by-user grouping,
    # coefficients written in by-user CSR order
    from truth_recommendation_gnn_b200.graph import CSR
    bu = ls.by_user
    eid_long = bu.eid.long()
    neg_by_user = CSR(bu.rowptr, neg.index_select(0, eid_long).int(), bu.eid, bu.n_rows, bu.n_cols)
    l_pos, c_pos_u, g_u = Fn.edge_anchor_loss(bu, ue, pe, c["Ee"], 1, ls.wbar, True, None, coef_in_csr_order=True)
    l_neg, c_neg_u, g_u = Fn.edge_anchor_loss(neg_by_user, ue, pe, c["Ee"], 0, ls.wbar, True, g_u, coef_in_csr_order=True)
    loss = l_pos + l_neg
    inv = torch.empty_like(eid_long)
    inv[eid_long] = torch.arange(c["Ee"], device=dev)          # original edge -> by-user CSR position
    c_pos, c_neg = c_pos_u[inv], c_neg_u[inv]
    # the two-row C-ABI form (trg_edge_bce_fwd) must agree with it
    loss2, c_pos2, c_neg2, g_u2 = Fn.edge_bce_fwd(ls, ue, pe, neg, want_grad=True)
    assert abs(float(loss2) - float(loss)) <= TOL_F32 * abs(float(loss))
    assert_close(g_u2, g_u, TOL_F32, "dL/du, two-row vs single-row form")
    del c_pos2, c_neg2, g_u2, c_pos_u, c_neg_u, inv, neg_by_user
    sp_pos = torch.zeros((), dtype=torch.float64, device=dev)
    sp_neg = torch.zeros((), dtype=torch.float64, device=dev)
    w_sum = torch.zeros((), dtype=torch.float64, device=dev)
    for a in range(0, c["Ee"], 4_000_000):
        b = min(a + 4_000_000, c["Ee"])
        uu = ue[pos_u[a:b]]
        sp_pos += F.softplus(-(uu * pe[pos_p[a:b]]).sum(1)).double().sum()       # BCEWithLogits(x, 1) = softplus(-x)
        sp_neg += F.softplus((uu * pe[neg[a:b]]).sum(1)).double().sum()
        w_sum += g.interaction_type_tensor[pos_p[a:b] + U].double().sum()
    e = float(c["Ee"])
    loss_ref = float(w_sum / e * sp_pos / e + sp_neg / e)
    assert abs(float(loss) - loss_ref) <= TOL_F32 * abs(loss_ref), (float(loss), loss_ref)
    # --- 100k sampled edges: dloss/dscore coefficients recomputed on the CPU from the embeddings ---
    sel = torch.randint(0, c["Ee"], (100_000,), generator=gen)
    seld = sel.to(dev)
    uu = ue[pos_u[seld]].cpu()
    s_pos = (uu * pe[pos_p[seld]].cpu()).sum(1).double()
    s_neg = (uu * pe[neg[seld]].cpu()).sum(1).double()
    wbar = float(w_sum / e)
    assert_close_elementwise(c_pos[seld].cpu(), -wbar * torch.sigmoid(-s_pos) / e, 2 * TOL_F32, "c_pos sample")
    assert_close_elementwise(c_neg[seld].cpu(), torch.sigmoid(s_neg) / e, 2 * TOL_F32, "c_neg sample")


@pytest.mark.timeout(1500)
def test_config5_full_catalogue_integer_exact(dev):
    """4096 queries x 50M posts, K = 100, bf16 (the benchmarked sweep) on integer-valued inputs: ids and
    values equal an exact ranking built from int64 keys (score, then lower id first)."""
    B, P, H, K = 4096, 50_000_000, 128, 100
    gen = torch.Generator(device=dev).manual_seed(11)
    q = torch.randint(0, 4, (B, H), generator=gen, device=dev).to(torch.bfloat16)
    cat = torch.randint(0, 4, (P, H), generator=gen, device=dev, dtype=torch.int8)
    dead = torch.rand(P, generator=gen, device=dev) < 0.05                     # all-zero rows: ties at score 0
    cat[dead] = 0
    cat = cat.to(torch.bfloat16)
    vals, ids = trg.score_topk(q, cat, K)
    assert vals.shape == (B, K) and ids.dtype == torch.int64
    # exact reference: key = score * 2^26 + (2^26 - 1 - id) is unique per post, so topk on keys IS the
    # canonical order; scores <= 9 * 128 are exact in fp32
    shift = 1 << 26
    best = torch.full((B, K), -1, dtype=torch.int64, device=dev)
    chunk = 500_000
    for a in range(0, P, chunk):
        b = min(a + chunk, P)
        sc = (q.float() @ cat[a:b].float().t()).long()
        key = sc * shift + (shift - 1 - torch.arange(a, b, device=dev))[None, :]
        best = torch.topk(torch.cat([best, key], dim=1), K, dim=1)[0]
        del sc, key
    ref_vals = (best // shift).float()
    ref_ids = shift - 1 - (best % shift)
    assert torch.equal(ids, ref_ids)
    assert torch.equal(vals, ref_vals)
