"""Row-partitioned collectives with autograd (torch.distributed: NCCL on the B200 box, gloo in the
CPU tests of the host logic)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def all_gather_rows_raw(x_local: torch.Tensor) -> torch.Tensor:
    world = dist.get_world_size()
    out = torch.empty(world * x_local.size(0), *x_local.shape[1:], dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, x_local.contiguous())
    return out


def reduce_scatter_rows_raw(g_full: torch.Tensor) -> torch.Tensor:
    world, rank = dist.get_world_size(), dist.get_rank()
    chunk = g_full.size(0) // world
    g_full = g_full.contiguous()
    if dist.get_backend() == "gloo":      # gloo has no reduce_scatter: all_reduce + slice
        g_full = g_full.clone()
        dist.all_reduce(g_full)
        return g_full[rank * chunk:(rank + 1) * chunk].clone()
    out = torch.empty(chunk, *g_full.shape[1:], dtype=g_full.dtype, device=g_full.device)
    dist.reduce_scatter_tensor(out, g_full, op=dist.ReduceOp.SUM)
    return out


class AllGatherRows(torch.autograd.Function):
    """[chunk, F] per rank -> [world*chunk, F]; backward = reduce-scatter (sum) of the partials."""

    @staticmethod
    def forward(ctx, x_local):
        return all_gather_rows_raw(x_local)

    @staticmethod
    def backward(ctx, g_full):
        if g_full.dtype == torch.bfloat16:      # cross-rank add in fp32, one rounding at the owner
            return reduce_scatter_rows_raw(g_full.float()).to(g_full.dtype)
        return reduce_scatter_rows_raw(g_full)


def all_gather_rows(x_local):
    if x_local.requires_grad:
        return AllGatherRows.apply(x_local)
    return all_gather_rows_raw(x_local)


class PushMeanAggFn(torch.autograd.Function):
    """Mean aggregation of a relation partitioned by SOURCE ("push"): this rank sums its own source
    rows into partial sums for ALL destinations, the partials are reduce-scattered to the destination
    owners and divided by the global in-degree.  Used for post -> user: the user table is 5x smaller
    than the post table, so moving [U, H] partial sums costs 5x less NVLink traffic than all-gathering
    [P, H] source rows.  Backward: all-gather the (1/deg-scaled) destination gradient, then a
    transposed gather over the local sources."""

    @staticmethod
    def forward(ctx, x_local, prel, gather_sum, grad_prescaled):
        # bf16 tables: fp32 partial sums cross NVLink and the mean is rounded once, at the owner (like the
        # single-GPU kernel: fp32 accumulate, one rounding)
        xfer = torch.float32 if x_local.dtype == torch.bfloat16 else None
        part = gather_sum(prel.rel, "fwd", x_local, xfer)          # [n_dst_pad, F]
        local = reduce_scatter_rows_raw(part)                      # owned destination rows
        ctx.prel, ctx.gather_sum, ctx.grad_prescaled = prel, gather_sum, grad_prescaled
        return (local * prel.inv_deg.to(local.dtype)[:, None]).to(x_local.dtype)

    @staticmethod
    def backward(ctx, g_mean):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        prel = ctx.prel
        if not ctx.grad_prescaled:
            g_mean = g_mean * prel.inv_deg.to(g_mean.dtype)[:, None]
        g_full = all_gather_rows_raw(g_mean.contiguous())
        return ctx.gather_sum(prel.rel, "bwd", g_full), None, None, None
