"""Oracle-backed compute ops injected by the CPU (gloo) tests of the multi-GPU HOST logic.
Test infrastructure: the product default is the CUDA kernels (nn.CUDA_OPS / dist.CUDA_LOSS_OPS)."""
import torch
import torch.nn.functional as F

from oracle import sage as osage
from oracle import topk as otopk


class OracleOps:
    @staticmethod
    def gather_sum(rel, which, x):
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


class OracleLossOps:
    @staticmethod
    def link_loss(user_local, post_full, shard, neg_local):
        pu, pp = shard.train_local[0], shard.train_local[1]
        pos = (user_local[pu] * post_full[pp]).sum(1)
        neg = (user_local[pu] * post_full[neg_local]).sum(1)
        e = float(shard.n_pos_global)
        return shard.wbar[0] * F.softplus(-pos).sum() / e + F.softplus(neg).sum() / e


def score_topk(q, cat, k, id_offset=0):
    return otopk.score_topk(q, cat, k, id_offset)


def merge(vals, ids, n_lists, k):
    k_in = vals.size(1) // n_lists
    return otopk.merge_topk([vals[:, i * k_in:(i + 1) * k_in] for i in range(n_lists)],
                            [ids[:, i * k_in:(i + 1) * k_in] for i in range(n_lists)], k)
