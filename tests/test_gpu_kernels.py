"""GPU parity tests, kernel level: every C-ABI entry point against the CPU oracle on the same
seeded inputs.  Bit-exact for integer / index work and for the fp32 CSR-order mean; 1e-5
relative for other fp32 results; 1e-2 for bf16 storage (fp32 accumulate)."""
import pytest
import torch

import truth_recommendation_gnn_b200 as trg
from oracle import cint
from oracle import csr as ocsr
from oracle import sage as osage
from oracle import topk as otopk
from tests.util import TOL_BF16, TOL_F32, assert_close
from truth_recommendation_gnn_b200 import functional as Fn
from truth_recommendation_gnn_b200 import synth

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------- K0
@pytest.mark.parametrize("n_src,n_dst,e,skew", [
    (50, 70, 600, False), (1000, 3000, 50_000, False), (300, 200, 40_000, True),
    (5, 1, 1000, False),            # a single destination: one very long row
    (70_000, 300_000, 400_000, False),  # 3 radix passes (19 bits), many empty rows
    (4, 4, 0, False),               # E = 0
    (0, 3, 0, False),               # N_src = 0 (inference.py:412-419)
    (10, 257, 5000, False),         # 2 passes, second one nearly empty
])
def test_csr_build_bit_exact(dev, n_src, n_dst, e, skew):
    g = torch.Generator().manual_seed(e + n_dst)
    src = torch.randint(0, max(n_src, 1), (e,), generator=g)
    if skew:
        dst = (n_dst * torch.rand(e, generator=g, dtype=torch.float64).pow(3)).long().clamp_(0, n_dst - 1)
    else:
        dst = torch.randint(0, max(n_dst, 1), (e,), generator=g)
    ei = torch.stack([src, dst])
    rp, col, eid = ocsr.csr_by_dst(ei, n_dst)
    c = trg.build_csr(ei[0].to(dev), ei[1].to(dev), n_dst, n_src)
    assert torch.equal(c.rowptr.cpu().long(), rp)
    assert torch.equal(c.eid.cpu().long(), eid)
    assert torch.equal(c.col.cpu().long(), col)
    if e:
        rp2, col2, eid2 = cint.csr_by_dst(ei, n_dst)   # plain-C oracle agrees too
        assert torch.equal(c.eid.cpu().long(), eid2) and torch.equal(c.rowptr.cpu().long(), rp2)
    # transpose = same routine with roles swapped
    rpt, colt, eidt = ocsr.csr_by_src(ei, n_src)
    ct = trg.build_csr(ei[1].to(dev), ei[0].to(dev), n_src, n_dst)
    assert torch.equal(ct.rowptr.cpu().long(), rpt) and torch.equal(ct.col.cpu().long(), colt)
    assert torch.equal(ct.eid.cpu().long(), eidt)


def test_csr_build_rejects_out_of_range(dev):
    ei = torch.tensor([[0, 1], [0, 5]], device=dev)
    with pytest.raises(IndexError):
        trg.build_csr(ei[0], ei[1], 3, 2)


def test_csr_is_deterministic(dev):
    g = synth.synth_graph(2000, 5000, 60_000, 10_000, 8, seed=9, skew=True).to(dev)
    ei = g.edge_index_dict[synth.REL_ENGAGE]
    a = trg.build_csr(ei[0], ei[1], 5000, 2000)
    b = trg.build_csr(ei[0], ei[1], 5000, 2000)
    assert torch.equal(a.eid, b.eid) and torch.equal(a.col, b.col) and torch.equal(a.rowptr, b.rowptr)


# ---------------------------------------------------------------------------------------- K1
def _agg_case(n_src, n_dst, e, f, seed, skew=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_src, f, generator=g)
    src = torch.randint(0, max(n_src, 1), (e,), generator=g)
    if skew:
        dst = (n_dst * torch.rand(e, generator=g, dtype=torch.float64).pow(3)).long().clamp_(0, n_dst - 1)
    else:
        dst = torch.randint(0, max(n_dst, 1), (e,), generator=g)
    return x, torch.stack([src, dst])


@pytest.mark.parametrize("f", [4, 8, 16, 32, 64, 96, 128, 256, 512])
def test_agg_fwd_fp32_bit_exact(dev, f):
    torch.set_num_threads(1)
    x, ei = _agg_case(500, 700, 6000, f, f)
    mean, cnt = osage.scatter_mean(x.index_select(0, ei[0]), ei[1], 700)
    rel = trg.RelationGraph(ei.to(dev), 500, 700)
    out, inv_deg = Fn.sage_agg_fwd(rel.fwd, x.to(dev))
    assert torch.equal(out.cpu(), mean), f"F={f}: max diff {(out.cpu() - mean).abs().max()}"
    assert torch.equal(inv_deg.cpu(), 1.0 / cnt)


def test_agg_fwd_edge_cases(dev):
    # E = 0, N_src = 0 -> all-zero means (inference.py:412-419)
    rel = trg.RelationGraph(torch.empty(2, 0, dtype=torch.long, device=dev), 0, 3)
    out, inv = Fn.sage_agg_fwd(rel.fwd, torch.empty(0, 64, device=dev))
    assert out.shape == (3, 64) and float(out.abs().sum()) == 0 and inv.tolist() == [1, 1, 1]
    # N_dst = 0
    rel = trg.RelationGraph(torch.empty(2, 0, dtype=torch.long, device=dev), 3, 0)
    out, _ = Fn.sage_agg_fwd(rel.fwd, torch.zeros(3, 64, device=dev))
    assert out.shape == (0, 64)
    # a hub destination, duplicates, isolated rows
    torch.set_num_threads(1)
    x, ei = _agg_case(300, 50, 20_000, 64, 3, skew=True)
    ei = torch.cat([ei, ei[:, :500]], dim=1)
    mean, cnt = osage.scatter_mean(x.index_select(0, ei[0]), ei[1], 50)
    out, _ = Fn.sage_agg_fwd(trg.RelationGraph(ei.to(dev), 300, 50).fwd, x.to(dev))
    from truth_recommendation_gnn_b200.graph import LONG_ROW_THRESHOLD
    deg = torch.bincount(ei[1], minlength=50)
    short = deg <= LONG_ROW_THRESHOLD
    assert short.any() and (~short).any()
    assert torch.equal(out.cpu()[short], mean[short])            # sequential CSR-order sum: bit-exact
    assert_close(out.cpu()[~short], mean[~short], TOL_F32, "hub rows (split into slices)")


@pytest.mark.parametrize("f,dtype,tol", [(128, torch.float32, TOL_F32), (64, torch.float32, TOL_F32),
                                         (128, torch.bfloat16, TOL_BF16)])
def test_long_row_splitting(dev, f, dtype, tol, monkeypatch):
    """Hub rows (power-law degrees) are cut into slices summed by a second kernel: same values to
    fp32 rounding, deterministic, forward / backward / weighted sum."""
    from truth_recommendation_gnn_b200 import graph as G
    monkeypatch.setattr(G, "LONG_ROW_THRESHOLD", 500)
    x, ei = _agg_case(3000, 400, 60_000, f, 77, skew=True)          # row 0 collects ~13% of the edges
    x = x.to(dtype).float()
    rel = trg.RelationGraph(ei.to(dev), 3000, 400)
    lr = rel.fwd.long_rows()
    assert lr is not None and lr.long_rows.numel() >= 3 and lr.n_slots > lr.long_rows.numel()
    xr = x.clone().requires_grad_(True)
    mean, cnt = osage.scatter_mean(xr.index_select(0, ei[0]), ei[1], 400)
    gup = torch.randn(400, f, generator=torch.Generator().manual_seed(2)).to(dtype).float()
    mean.backward(gup)
    xd = x.to(dev).to(dtype).requires_grad_(True)
    out = trg.sage_mean_aggregate(xd, rel)
    out.backward(gup.to(dev).to(dtype))
    assert_close(out.float().cpu(), mean.detach(), tol, "split-row mean")
    assert_close(xd.grad.float().cpu(), xr.grad, tol, "split-row backward (transposed CSR has hub sources)")
    out2 = trg.sage_mean_aggregate(xd.detach(), rel)
    assert torch.equal(out2, out.detach())                           # deterministic
    coef = torch.randn(60_000, generator=torch.Generator().manual_seed(3))
    exp = torch.zeros(400, f, dtype=torch.float64).index_add_(0, ei[1], coef.double()[:, None] * x.double()[ei[0]])
    ws = Fn.gather_wsum(rel.fwd, coef.to(dev), x.to(dev).to(dtype))
    assert_close(ws.float().cpu(), exp, tol, "split-row weighted sum")
    Fn.gather_wsum(rel.fwd, coef.to(dev), x.to(dev).to(dtype), out=ws, accumulate=True)
    assert_close(ws.float().cpu(), 2 * exp, 2 * tol, "split-row weighted sum, accumulate")


@pytest.mark.parametrize("f", [64, 128, 256])
def test_agg_fwd_bf16(dev, f):
    x, ei = _agg_case(400, 600, 5000, f, 100 + f)
    xb = x.bfloat16()
    mean, _ = osage.scatter_mean(xb.float().index_select(0, ei[0]), ei[1], 600)
    out, _ = Fn.sage_agg_fwd(trg.RelationGraph(ei.to(dev), 400, 600).fwd, xb.to(dev))
    assert out.dtype == torch.bfloat16
    # fp32 accumulate then one bf16 rounding: error <= 2^-8 relative per element
    assert_close(out.float().cpu(), mean, TOL_BF16, f"bf16 agg F={f}")
    assert torch.equal(out.cpu(), mean.bfloat16())  # same fp32 sum, same rounding


# ---------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("f,dtype,tol", [(64, torch.float32, TOL_F32), (128, torch.float32, TOL_F32),
                                         (32, torch.float32, TOL_F32), (128, torch.bfloat16, TOL_BF16)])
def test_agg_bwd_vs_autograd(dev, f, dtype, tol):
    x, ei = _agg_case(300, 450, 4000, f, 200 + f, skew=True)
    x = x.to(dtype).float()
    g = torch.randn(450, f, generator=torch.Generator().manual_seed(1)).to(dtype).float()
    xr = x.clone().requires_grad_(True)
    mean, _ = osage.scatter_mean(xr.index_select(0, ei[0]), ei[1], 450)
    mean.backward(g)
    rel = trg.RelationGraph(ei.to(dev), 300, 450)
    xd = x.to(dev).to(dtype).requires_grad_(True)
    out = trg.sage_mean_aggregate(xd, rel)
    out.backward(g.to(dev).to(dtype))
    assert_close(xd.grad.float().cpu(), xr.grad, tol, "agg bwd")


def test_agg_bwd_skipped_for_leaf_features(dev):
    # layer-1 inputs are constant features -> no transposed CSR is ever built
    x, ei = _agg_case(50, 60, 300, 64, 5)
    rel = trg.RelationGraph(ei.to(dev), 50, 60)
    w = torch.randn(64, 64, device=dev, requires_grad=True)
    (trg.sage_mean_aggregate(x.to(dev), rel) @ w).sum().backward()
    assert rel._bwd is None and w.grad is not None


# ------------------------------------------------------------------------------------ wsum
def test_gather_wsum(dev):
    x, ei = _agg_case(200, 300, 3000, 64, 17)
    coef = torch.randn(3000, generator=torch.Generator().manual_seed(2))
    exp = torch.zeros(300, 64, dtype=torch.float64).index_add_(
        0, ei[1], coef.double()[:, None] * x.double()[ei[0]])
    csr = trg.RelationGraph(ei.to(dev), 200, 300).fwd
    scale = torch.tensor([0.5], device=dev)
    out = Fn.gather_wsum(csr, coef.to(dev), x.to(dev), scale=scale)
    assert_close(out.cpu(), 0.5 * exp, TOL_F32, "wsum")
    Fn.gather_wsum(csr, coef.to(dev), x.to(dev), scale=None, out=out, accumulate=True)
    assert_close(out.cpu(), 1.5 * exp, TOL_F32, "wsum accumulate")


@pytest.mark.parametrize("f,dtype,tol,long_rows", [(128, torch.float32, TOL_F32, False), (64, torch.float32, TOL_F32, False),
                                                   (16, torch.float32, TOL_F32, False), (8, torch.float32, TOL_F32, False),
                                                   (128, torch.float32, TOL_F32, True), (128, torch.bfloat16, TOL_BF16, False)])
def test_gather_epilogue_accumulate_and_relu_gate(dev, f, dtype, tol, long_rows, monkeypatch):
    """accumulate + relu_of epilogues of K2 / wsum (the fused ReLU backward and gradient
    accumulation of fused_step): out = where(act > 0, base + gather_sum, 0)."""
    if long_rows:
        from truth_recommendation_gnn_b200 import graph as G
        monkeypatch.setattr(G, "LONG_ROW_THRESHOLD", 64)
    x, ei = _agg_case(300, 200, 6000, f, 77 + f, skew=True)
    gen = torch.Generator().manual_seed(9)
    x = x.to(dtype).float()
    base = torch.randn(200, f, generator=gen).to(dtype).float()
    act = torch.relu(torch.randn(200, f, generator=gen)).to(dtype).float()
    coef = torch.randn(6000, generator=gen)
    csr = trg.RelationGraph(ei.to(dev), 300, 200).fwd
    s_plain = torch.zeros(200, f, dtype=torch.float64).index_add_(0, ei[1], x.double()[ei[0]])
    s_coef = torch.zeros(200, f, dtype=torch.float64).index_add_(0, ei[1], coef.double()[:, None] * x.double()[ei[0]])
    xd, actd = x.to(dev).to(dtype), act.to(dev).to(dtype)
    for what, run, ssum in (
            ("agg_bwd", lambda o, **kw: Fn.sage_agg_bwd(csr, None, xd, out=o, **kw), s_plain),
            ("wsum", lambda o, **kw: Fn.gather_wsum(csr, coef.to(dev), xd, out=o, **kw), s_coef)):
        out = run(base.to(dev).to(dtype).clone(), accumulate=True)
        assert_close(out.float().cpu(), base.double() + ssum, tol, what + " accumulate")
        out = run(base.to(dev).to(dtype).clone(), accumulate=True, relu_of=actd)
        exp = torch.where(act > 0, base.double() + ssum, torch.zeros((), dtype=torch.float64))
        assert_close(out.float().cpu(), exp, tol, what + " accumulate + gate")
        assert bool((out.float().cpu()[act <= 0] == 0).all())
        out = run(None, relu_of=actd)
        exp = torch.where(act > 0, ssum, torch.zeros((), dtype=torch.float64))
        assert_close(out.float().cpu(), exp, tol, what + " gate only")


# ---------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("h,dtype,tol", [(64, torch.float32, TOL_F32), (128, torch.float32, TOL_F32),
                                         (16, torch.float32, TOL_F32), (256, torch.float32, TOL_F32),
                                         (128, torch.bfloat16, TOL_BF16),
                                         (256, torch.bfloat16, TOL_BF16)])   # 512-byte bf16 rows: cp.async ring form
def test_link_bce_fwd_bwd(dev, h, dtype, tol):
    U, P, E = 300, 500, 4000
    g = torch.Generator().manual_seed(h)
    u = torch.relu(torch.randn(U, h, generator=g)).to(dtype).float()
    p = torch.relu(torch.randn(P, h, generator=g)).to(dtype).float()
    tei = torch.stack([torch.randint(0, U, (E,), generator=g), torch.randint(0, P, (E,), generator=g)])
    tei = torch.cat([tei, tei[:, :50]], dim=1)          # duplicate positives
    neg = torch.randint(0, P, (tei.size(1),), generator=g)
    w = torch.zeros(U + P)
    w[U:] = torch.where(torch.rand(P, generator=g) < 0.25, 3.0, 1.0)
    ur, pr = u.clone().requires_grad_(True), p.clone().requires_grad_(True)
    loss = osage.link_loss(ur, pr, tei[0], tei[1], neg, w, U)
    (2.0 * loss).backward()
    ud, pd = u.to(dev).to(dtype).requires_grad_(True), p.to(dev).to(dtype).requires_grad_(True)
    lg = trg.link_bce_loss(ud, pd, tei.to(dev), neg.to(dev), w.to(dev), U)
    (2.0 * lg).backward()
    assert abs(float(lg) - float(loss)) <= tol * abs(float(loss)), (float(lg), float(loss))
    assert_close(ud.grad.float().cpu(), ur.grad, tol, "dL/du")
    assert_close(pd.grad.float().cpu(), pr.grad, tol, "dL/dp")
    # forward only (eval): no grads requested
    with torch.no_grad():
        l2 = trg.link_bce_loss(ud, pd, tei.to(dev), neg.to(dev), w.to(dev), U)
    assert float(l2) == float(lg)


def test_link_bce_deterministic(dev):
    U, P, E, h = 200, 300, 5000, 64
    g = torch.Generator().manual_seed(0)
    u = torch.randn(U, h, generator=g).to(dev).requires_grad_(True)
    p = torch.randn(P, h, generator=g).to(dev).requires_grad_(True)
    tei = torch.stack([torch.randint(0, U, (E,), generator=g), torch.randint(0, P, (E,), generator=g)]).to(dev)
    neg = torch.randint(0, P, (E,), generator=g).to(dev)
    w = torch.ones(U + P, device=dev)
    res = []
    for _ in range(2):
        u.grad = p.grad = None
        l = trg.link_bce_loss(u, p, tei, neg, w, U)
        l.backward()
        res.append((l.clone(), u.grad.clone(), p.grad.clone()))
    assert all(torch.equal(a, b) for a, b in zip(*res))   # atomic-free: bitwise repeatable


# ---------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n,h,ks,dtype,tol", [
    (1000, 64, (64, 64, 64, 64), torch.float32, TOL_F32),      # tcgen05 3xTF32, 4 terms, tail tile
    (777, 128, (128, 128), torch.float32, TOL_F32),
    (128 * 200 + 5, 128, (128, 128, 128), torch.float32, TOL_F32),   # > 148 tiles: persistent loop wraps
    (128 * 1500, 128, (128, 128), torch.float32, TOL_F32),            # ~10 tiles per CTA: every ring wraps
    (4000, 256, (256, 256), torch.float32, TOL_F32),
    (1, 64, (64, 64, 64, 64), torch.float32, TOL_F32),
    (300, 32, (16, 48), torch.float32, TOL_F32),                # not tensor-core shaped -> FMA kernel
    (513, 128, (128, 128, 128), torch.bfloat16, TOL_BF16),      # tcgen05 kind::f16
    (20_000, 256, (256, 256), torch.bfloat16, TOL_BF16),
    (900, 64, (64, 128), torch.bfloat16, TOL_BF16),
])
def test_proj_fwd(dev, n, h, ks, dtype, tol):
    g = torch.Generator().manual_seed(n + h)
    alphas = [1.0, 1.0, 0.75, 0.75][:len(ks)]
    terms = [(torch.randn(n, k, generator=g).to(dtype), (torch.randn(h, k, generator=g) / k ** 0.5).to(dtype), a)
             for k, a in zip(ks, alphas)]
    bias = torch.randn(h, generator=g)
    exp = sum(a * (A.double() @ W.double().t()) for A, W, a in terms) + bias.double()
    for relu in (False, True):
        out = Fn.sage_proj_fwd([(A.to(dev), W.to(dev), a) for A, W, a in terms], bias.to(dev), relu)
        e = torch.relu(exp) if relu else exp
        assert_close(out.float().cpu(), e, tol, f"proj relu={relu}")


def test_proj_fwd_kblock_signature_many_tiles(dev):
    """Regression for a write-after-read race on the activation landing ring (a refill TMA overtook a
    late shared-memory read once the ring ran ahead, i.e. only from a CTA's second tile on): k-block j of
    the concatenated K contributes exactly 2^j to every output, so any k-block that is dropped, doubled
    or taken from another pipeline stage changes the integer result."""
    n, h, nt = 128 * 1200, 128, 2
    rows = torch.arange(n, device=dev)
    A = [torch.zeros(n, 128, device=dev) for _ in range(nt)]
    for t in range(nt):
        for j in range(4):
            A[t][rows, 32 * j + rows % 32] = float(2 ** (t * 4 + j))
    W = [torch.ones(h, 128, device=dev) for _ in range(nt)]
    for _ in range(3):
        out = Fn.sage_proj_fwd([(a, w, 1.0) for a, w in zip(A, W)], None, False)
        assert bool((out == 255.0).all()), torch.unique(out).tolist()[:8]
        outs = Fn.sage_proj_bwd_input(A[0], [(w, 1.0, None) for w in W])
        assert all(bool((o == 15.0).all()) for o in outs)


@pytest.mark.parametrize("n,h,ks,dtype,tol", [
    (1000, 128, (128, 128, 128), torch.float32, TOL_F32),
    (128 * 160 + 77, 64, (64, 64), torch.float32, TOL_F32),
    (500, 256, (256,), torch.float32, TOL_F32),
    (300, 32, (16, 16), torch.float32, TOL_F32),                # FMA kernel path
    (2000, 128, (128, 128), torch.bfloat16, TOL_BF16),
])
def test_proj_bwd_input(dev, n, h, ks, dtype, tol):
    g = torch.Generator().manual_seed(n * 3 + h)
    dz = torch.randn(n, h, generator=g).to(dtype)
    ws = [(torch.randn(h, k, generator=g) / h ** 0.5).to(dtype) for k in ks]
    alphas = [1.0, 0.75, 1.0][:len(ks)]
    rs = [torch.rand(n, generator=g) if i % 2 == 0 else None for i in range(len(ks))]
    outs = Fn.sage_proj_bwd_input(dz.to(dev), [(w.to(dev), a, r.to(dev) if r is not None else None)
                                               for w, a, r in zip(ws, alphas, rs)])
    for o, w, a, r in zip(outs, ws, alphas, rs):
        exp = a * (dz.double() @ w.double())
        if r is not None:
            exp = exp * r.double()[:, None]
        assert_close(o.float().cpu(), exp, tol, "proj bwd input")


@pytest.mark.parametrize("n,h,ks,dtype,tol", [
    (1000, 128, (128, 128, 128), torch.float32, TOL_F32),      # tcgen05 MN-major 3xTF32
    (50_000, 128, (128, 128), torch.float32, TOL_F32),
    (7, 128, (128,), torch.float32, TOL_F32),                   # fewer rows than one k-block
    (3000, 64, (64, 64, 64), torch.float32, TOL_F32),           # hidden < 128: TMA zero fill
    (3000, 256, (256, 256), torch.float32, TOL_F32),            # two 128-wide blocks of hidden
    (2000, 128, (64, 256), torch.float32, TOL_F32),             # mixed widths
    (300, 32, (16, 48), torch.float32, TOL_F32),                # generic FMA path
    (5000, 128, (128, 128, 128), torch.bfloat16, TOL_BF16),
    (5000, 256, (256, 256), torch.bfloat16, TOL_BF16),
])
def test_proj_bwd_weight(dev, n, h, ks, dtype, tol):
    g = torch.Generator().manual_seed(n * 7 + h)
    dz = torch.randn(n, h, generator=g).to(dtype)
    dz[torch.rand(n, h, generator=g) < 0.5] = 0            # post-ReLU-mask sparsity
    A = [torch.randn(n, k, generator=g).to(dtype) for k in ks]
    alphas = [1.0, 0.75, 1.0][:len(ks)]
    outs, db = Fn.sage_proj_bwd_weight(dz.to(dev), [(a.to(dev), al) for a, al in zip(A, alphas)], True)
    for o, a, al in zip(outs, A, alphas):
        exp = al * (dz.double().t() @ a.double())
        assert o.shape == exp.shape
        assert_close(o.float().cpu(), exp, tol, "dW")
    assert_close(db.cpu(), dz.double().sum(0), tol, "db")
    outs2, db2 = Fn.sage_proj_bwd_weight(dz.to(dev), [(a.to(dev), al) for a, al in zip(A, alphas)], True)
    assert all(torch.equal(x, y) for x, y in zip(outs, outs2)) and torch.equal(db, db2)   # deterministic


# ---------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("b,p,h,k", [(1, 20_000, 64, 10), (37, 5000, 64, 10), (64, 3000, 128, 100),
                                     (5, 7, 64, 10), (3, 1000, 16, 3), (130, 2500, 256, 128)])
def test_score_topk_integer_exact(dev, b, p, h, k):
    """Small-integer embeddings make every dot product exact in fp32 whatever the summation
    order, so values AND ids (with massive ties) must equal the canonical oracle bit for bit."""
    g = torch.Generator().manual_seed(b * p)
    q = torch.randint(0, 4, (b, h), generator=g).float()
    cat = torch.randint(0, 3, (p, h), generator=g).float()
    cat[torch.rand(p, generator=g) < 0.05] = 0          # dead-ReLU rows: exact-zero scores
    vals, ids = otopk.score_topk(q, cat, k)
    gv, gi = trg.score_topk(q.to(dev), cat.to(dev), k)
    assert torch.equal(gv.cpu(), vals) and torch.equal(gi.cpu(), ids)
    assert torch.equal(gv.cpu(), torch.topk(q @ cat.t(), min(k, p))[0])   # literal torch.topk values
    gv2, gi2 = trg.score_topk(q.to(dev), cat.to(dev), k, id_offset=1000)
    assert torch.equal(gi2.cpu(), ids + 1000)


@pytest.mark.parametrize("b,p,h,k", [(1, 20_000, 64, 10), (300, 50_000, 128, 100), (129, 7001, 256, 100), (70, 9000, 192, 100),
                                     (4096, 3000, 128, 100), (5, 7, 64, 10), (64, 200_000, 128, 100)])
def test_score_topk_bf16_tensor_core_integer_exact(dev, b, p, h, k):
    """bf16 tcgen05 path: small-integer embeddings are exact in bf16 and their dot products exact in
    the fp32 accumulator, so values and ids (with massive ties) must match the oracle bit for bit."""
    g = torch.Generator().manual_seed(b * p + h)
    q = torch.randint(0, 4, (b, h), generator=g).float()
    cat = torch.randint(0, 3, (p, h), generator=g).float()
    cat[torch.rand(p, generator=g) < 0.05] = 0
    vals, ids = otopk.score_topk(q, cat, k)
    gv, gi = trg.score_topk(q.to(dev).bfloat16(), cat.to(dev).bfloat16(), k, id_offset=7)
    assert torch.equal(gv.cpu(), vals) and torch.equal(gi.cpu(), ids + 7)


def test_score_topk_bf16_large_catalogue(dev):
    """4.2M posts (tens of thousands of tiles per split, thresholds shared across splits): bit-exact, ties
    included.  (A pre-warm of the thresholds from the top-K of the first 4096 posts was measured on config 5:
    exact, but 84.18 vs 84.19 ms -- the cold start of the lists is not what costs time -- and removed.)"""
    b, p, h, k = 9, (1 << 22) + 12_345, 64, 20
    g = torch.Generator().manual_seed(99)
    q = torch.randint(0, 4, (b, h), generator=g).float()
    cat = torch.randint(0, 3, (p, h), generator=g).float()
    cat[torch.rand(p, generator=g) < 0.05] = 0
    vals, ids = otopk.score_topk(q, cat, k)
    gv, gi = trg.score_topk(q.to(dev).bfloat16(), cat.to(dev).bfloat16(), k, id_offset=3)
    assert torch.equal(gv.cpu(), vals) and torch.equal(gi.cpu(), ids + 3)


def test_score_topk_bf16_random(dev):
    q, cat = synth.synth_queries(200, 100_000, 128, zero_frac=0.05)
    qb, cb = q.bfloat16(), cat.bfloat16()
    scores = qb.float() @ cb.float().t()                 # fp32 oracle on bf16-rounded inputs
    vals, ids = otopk.topk_canonical(scores, 100)
    gv, gi = trg.score_topk(qb.to(dev), cb.to(dev), 100)
    assert_close(gv.cpu(), vals, 1e-5, "bf16 top-k values (fp32 accumulate)")
    s_at = scores.gather(1, gi.cpu())
    assert_close(s_at, gv.cpu(), 1e-5, "scores at returned ids")
    assert (gi.cpu() == ids).float().mean() > 0.99       # only near-ties may swap


def test_score_topk_float_and_sharded(dev):
    q, cat = synth.synth_queries(50, 30_000, 64, zero_frac=0.05)
    scores = q @ cat.t()
    vals, ids = otopk.topk_canonical(scores, 10)
    gv, gi = trg.score_topk(q.to(dev), cat.to(dev), 10)
    assert_close(gv.cpu(), vals, TOL_F32, "top-k values")
    # ids: exact wherever the oracle's neighbouring scores are separated by more than rounding
    s_at = scores.gather(1, gi.cpu())
    assert_close(s_at, gv.cpu(), TOL_F32, "scores at returned ids")
    srt = torch.sort(scores, dim=1, descending=True)[0][:, :11]
    gap_ok = ((srt[:, :-1] - srt[:, 1:]) > 1e-4 * srt[:, :1]).all(dim=1)
    assert gap_ok.sum() > 10
    assert torch.equal(gi.cpu()[gap_ok], ids[gap_ok])
    # sharded catalogue + merge == unsharded
    bounds = [(0, 9000), (9000, 21_000), (21_000, 30_000)]
    parts = [trg.score_topk(q.to(dev), cat[a:b].to(dev).contiguous(), 10, id_offset=a) for a, b in bounds]
    mv, mi = trg.topk_merge(torch.cat([x[0] for x in parts], 1), torch.cat([x[1] for x in parts], 1), 3, 10)
    assert torch.equal(mv, gv) and torch.equal(mi, gi)


# ------------------------------------------------------------- fp32 sum rows / row finish (multi-GPU path)
@pytest.mark.parametrize("f,long_rows", [(128, False), (256, False), (64, False), (16, False), (128, True)])
def test_gather_fp32_output_rows_for_bf16_tables(dev, f, long_rows, monkeypatch):
    """out_dtype = fp32 with a bf16 table (the partial sums a multi-GPU run reduces across ranks): the rows
    are the UNROUNDED fp32 accumulations -- 1e-5 against an fp64 sum of the bf16 inputs, where the bf16
    output would only be good to 4e-3 -- including accumulate and the long-row combine."""
    if long_rows:
        from truth_recommendation_gnn_b200 import graph as G
        monkeypatch.setattr(G, "LONG_ROW_THRESHOLD", 64)
    x, ei = _agg_case(300, 200, 6000, f, 5 + f, skew=True)
    xb = x.bfloat16()
    gen = torch.Generator().manual_seed(3)
    coef = torch.randn(6000, generator=gen)
    csr = trg.RelationGraph(ei.to(dev), 300, 200).fwd
    s_plain = torch.zeros(200, f, dtype=torch.float64).index_add_(0, ei[1], xb.double()[ei[0]])
    s_coef = torch.zeros(200, f, dtype=torch.float64).index_add_(0, ei[1], coef.double()[:, None] * xb.double()[ei[0]])
    out = Fn.sage_agg_bwd(csr, None, xb.to(dev), out_dtype=torch.float32)
    assert out.dtype == torch.float32
    assert_close(out.cpu(), s_plain, TOL_F32, "fp32 sum rows of a bf16 table")
    Fn.gather_wsum(csr, coef.to(dev), xb.to(dev), out=out, accumulate=True)           # accumulates in fp32
    assert_close(out.cpu(), s_plain + s_coef, TOL_F32, "fp32 rows, accumulate")
    rounded = Fn.sage_agg_bwd(csr, None, xb.to(dev))
    assert rounded.dtype == torch.bfloat16
    assert_close(rounded.float().cpu(), s_plain, TOL_BF16, "bf16 rows")
    with pytest.raises(trg._lib.TrgError):
        Fn.sage_agg_bwd(csr, None, x.to(dev), out_dtype=torch.bfloat16)               # only bf16 -> fp32 exists


@pytest.mark.parametrize("dtype,in_f32", [(torch.float32, True), (torch.bfloat16, True), (torch.bfloat16, False)])
def test_rows_finish(dev, dtype, in_f32):
    """trg_rows_finish: out = gate(row_scale * in + add), fp32 arithmetic, one rounding."""
    gen = torch.Generator().manual_seed(1)
    n, f = 1000, 128
    x = torch.randn(n, f, generator=gen)
    x = x if in_f32 else x.bfloat16().float()
    add = torch.randn(n, f, generator=gen).to(dtype).float()
    act = torch.relu(torch.randn(n, f, generator=gen)).to(dtype).float()
    rs = torch.rand(n, generator=gen)
    xd = x.to(dev) if in_f32 else x.to(dev).to(dtype)
    tol = TOL_F32 if dtype == torch.float32 else 2.0 ** -8          # one bf16 rounding of the result
    out = Fn.rows_finish(xd, dtype, row_scale=rs.to(dev))
    assert out.dtype == dtype
    assert_close(out.float().cpu(), x.double() * rs.double()[:, None], tol, "row scale")
    out = Fn.rows_finish(xd, dtype, add=add.to(dev).to(dtype), relu_of=act.to(dev).to(dtype))
    exp = torch.where(act > 0, x.double() + add.double(), torch.zeros((), dtype=torch.float64))
    assert_close(out.float().cpu(), exp, tol, "add + gate")
    assert bool((out.float().cpu()[act <= 0] == 0).all())
    if dtype == torch.float32:
        assert torch.equal(out.cpu(), torch.where(act > 0, x + add, torch.zeros(())))    # exactly torch's fp32
    assert Fn.rows_finish(xd[:0], dtype).shape == (0, f)


@pytest.mark.parametrize("dtype,in_f32", [(torch.float32, True), (torch.bfloat16, True), (torch.bfloat16, False)])
@pytest.mark.parametrize("n_peers", [1, 2, 3, 4, 8])
def test_peer_reduce_rows(dev, dtype, in_f32, n_peers):
    """trg_peer_reduce_rows = reduce-scatter + trg_rows_finish in one kernel: the owned row range of G partial
    tables summed in rank order (bit-equal to the sequential fp32 sum), then 1/deg, local term, ReLU gate, one
    rounding.  The G tables live on this device here; over NVLink they are the peers' mapped buffers."""
    gen = torch.Generator().manual_seed(7 + n_peers)
    rows, f, row0, n = 900, 128, 300, 500
    parts = [torch.randn(rows, f, generator=gen) for _ in range(n_peers)]
    if not in_f32:
        parts = [p.bfloat16().float() for p in parts]
    add = torch.randn(n, f, generator=gen).to(dtype).float()
    act = torch.relu(torch.randn(n, f, generator=gen)).to(dtype).float()
    rs = torch.rand(n, generator=gen)
    pd = [p.to(dev) if in_f32 else p.to(dev).to(dtype) for p in parts]
    seq = parts[0][row0:row0 + n].clone()
    for p in parts[1:]:
        seq = seq + p[row0:row0 + n]                       # fp32, rank order
    out = Fn.peer_reduce_rows(pd, row0, n, dtype, row_scale=rs.to(dev))
    assert out.dtype == dtype and out.shape == (n, f)
    exp = seq * rs[:, None]
    if dtype == torch.float32:
        assert torch.equal(out.cpu(), exp)
    else:
        assert torch.equal(out.cpu(), exp.to(dtype))
    out = Fn.peer_reduce_rows(pd, row0, n, dtype, add=add.to(dev).to(dtype), relu_of=act.to(dev).to(dtype))
    exp = torch.where(act > 0, seq + add, torch.zeros(()))
    assert torch.equal(out.cpu(), exp.to(dtype))
    # same result as the two-step form it replaces
    two = Fn.rows_finish(torch.stack([p[row0:row0 + n].float() for p in pd]).sum(0) if n_peers > 2 else
                         (pd[0][row0:row0 + n].float() + (pd[1][row0:row0 + n].float() if n_peers == 2 else 0)),
                         dtype, add=add.to(dev).to(dtype), relu_of=act.to(dev).to(dtype))
    tol = TOL_F32 if dtype == torch.float32 else 2.0 ** -7
    assert_close(out.float().cpu(), two.float().cpu(), tol, "fused vs reduce + finish")
    assert Fn.peer_reduce_rows(pd, 0, 0, dtype).shape == (0, f)


def test_csr_build_aborts_on_out_of_range_per_step_ids(dev):
    """Per-step structures skip the host-side range check (no sync inside a step); K0's histogram pass then
    aborts the kernel on an id outside [0, n_key) instead of corrupting memory.  The abort poisons the CUDA
    context, so it is observed in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch, truth_recommendation_gnn_b200 as trg\n"
            "k = torch.tensor([0, 7, 2], device='cuda'); o = torch.tensor([0, 1, 1], device='cuda')\n"
            "trg.build_csr(o, k, 3, 2, validate=False, per_step=True)\n"
            "torch.cuda.synchronize()\nprint('NOT DETECTED')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)
    assert r.returncode != 0 and "NOT DETECTED" not in r.stdout, r.stdout + r.stderr
    assert "outside [0, 3)" in r.stdout + r.stderr
