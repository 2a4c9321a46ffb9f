"""Oracle-backed compute ops injected by the CPU (gloo) tests of the multi-GPU HOST logic.
Test infrastructure: the product default is the CUDA kernels (nn.CUDA_OPS / dist.CUDA_LOSS_OPS)."""
import torch
import torch.nn.functional as F

from oracle import sage as osage
from oracle import topk as otopk


class OracleOps:
    @staticmethod
    def gather_sum(rel, which, x, out_dtype=None):
        ei = rel.edge_index
        if which == "fwd":
            return torch.zeros(rel.n_dst, x.size(1), dtype=x.dtype).index_add_(0, ei[1], x[ei[0]])
        return torch.zeros(rel.n_src, x.size(1), dtype=x.dtype).index_add_(0, ei[0], x[ei[1]])

    @staticmethod
    def aggregate(x_src, rel):
        from truth_recommendation_gnn_b200.collectives import PushMeanAggFn
        from truth_recommendation_gnn_b200.graph import PushRelation
        if isinstance(rel, PushRelation):      # host logic under test; the oracle supplies the primitive
            return PushMeanAggFn.apply(x_src, rel, OracleOps.gather_sum, False)
        ei = rel.edge_index
        return osage.scatter_mean(x_src.index_select(0, ei[0]), ei[1], rel.n_dst)[0]

    @staticmethod
    def project(terms, bias, relu, scale_rels):
        out = sum(alpha * F.linear(a, w) for a, w, alpha in terms)
        if bias is not None:
            out = out + bias
        return torch.relu(out) if relu else out


class _Csr:
    def __init__(self, other, key, n_rows):
        self.other, self.key, self.n_rows = other, key, n_rows


class OracleLossOps:
    """Oracle primitives for dist.AnchoredLinkLossFn (post-owner loss partition)."""

    @staticmethod
    def csr(other, key, n_key, n_other, per_step=False):
        return _Csr(other, key, n_key)

    @staticmethod
    def anchor_loss(csr, post_local, user_full, n_edges, label, wbar, want, g_post):
        d = (post_local[csr.key] * user_full[csr.other]).sum(1)
        z = -d if label else d
        scale = float(wbar[0]) if label else 1.0
        loss = (scale * F.softplus(z).sum() / n_edges).reshape(1)
        if not want:
            return loss.detach(), None, None
        coef = (-scale if label else scale) * torch.sigmoid(z) / n_edges
        g = torch.zeros_like(post_local).index_add_(0, csr.key, coef[:, None] * user_full[csr.other])
        return loss.detach(), coef.detach(), (g if g_post is None else g_post + g).detach()

    @staticmethod
    def wsum(csr, coef, post_local, scale, out):
        g = torch.zeros(csr.n_rows, post_local.size(1), dtype=post_local.dtype)
        g.index_add_(0, csr.key, coef[:, None] * post_local[csr.other])
        g = g * scale
        return g if out is None else out + g


def score_topk(q, cat, k, id_offset=0):
    return otopk.score_topk(q, cat, k, id_offset)


def merge(vals, ids, n_lists, k):
    k_in = vals.size(1) // n_lists
    return otopk.merge_topk([vals[:, i * k_in:(i + 1) * k_in] for i in range(n_lists)],
                            [ids[:, i * k_in:(i + 1) * k_in] for i in range(n_lists)], k)


class OracleStepPrims:
    """Oracle primitives for dist_fused.loss_and_grads_sharded (the names of CudaStepPrims)."""

    @staticmethod
    def agg_mean(rel, x_src):
        ei = rel.edge_index
        mean, cnt = osage.scatter_mean(x_src.index_select(0, ei[0]), ei[1], rel.n_dst)
        if rel.inv_deg is None:
            rel.inv_deg = 1.0 / torch.bincount(ei[1], minlength=rel.n_dst).clamp(min=1).float()
        return mean

    @staticmethod
    def gather_sum(rel, which, x, out=None, accumulate=False, relu_of=None, out_dtype=None):
        g = OracleOps.gather_sum(rel, which, x)
        if out is not None and accumulate:
            g = out + g
        if relu_of is not None:
            g = torch.where(relu_of > 0, g, torch.zeros_like(g))
        return g

    @staticmethod
    def rows_finish(x, dtype, row_scale=None, add=None, relu_of=None):
        y = x.float()
        if row_scale is not None:
            y = y * row_scale[:, None]
        if add is not None:
            y = y + add.float()
        if relu_of is not None:
            y = torch.where(relu_of > 0, y, torch.zeros_like(y))
        return y.to(dtype)

    @staticmethod
    def proj_fwd(terms, bias, relu):
        return OracleOps.project(terms, bias, relu, None)

    @staticmethod
    def proj_bwd_weight(dz, terms, want_bias):
        return [alpha * (dz.t() @ a) for a, alpha in terms], (dz.sum(0) if want_bias else None)

    @staticmethod
    def proj_bwd_input(dz, terms):
        outs = []
        for w, alpha, rs in terms:
            d = alpha * (dz @ w)
            outs.append(d * rs[:, None] if rs is not None else d)
        return outs

    csr = staticmethod(OracleLossOps.csr)

    @staticmethod
    def select_negatives(shard, neg_p_global, capacity=None):
        return shard.local_negatives(neg_p_global)        # host logic under test: exact selection

    @staticmethod
    def anchor_loss(csr, anchor, gathered, n_edges, label, wbar, g_anchor, relu_gate):
        loss, coef, g = OracleLossOps.anchor_loss(csr, anchor, gathered, n_edges, label, wbar, True, g_anchor)
        if relu_gate:
            g = torch.where(anchor > 0, g, torch.zeros_like(g))
        return loss, coef, g

    @staticmethod
    def wsum(csr, coef, x, out=None, accumulate=False, out_dtype=None):
        return OracleLossOps.wsum(csr, coef, x, 1.0, out if accumulate else None)
