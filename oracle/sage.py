"""Pure-torch CPU restatement of the reference's message-passing model and train step.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``SAGEConv`` parity is *unpinned*: PyG is
not vendored by the reference and not installable here; this follows PyG 2.x's published
algorithm for exactly the options the reference uses.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

REL_DIRECT = ("post", "rev_engages", "user")   # train_gnn.py:142,179
REL_SOCIAL = ("user", "social", "user")        # train_gnn.py:122,183
REL_ENGAGE = ("user", "engages", "post")       # train_gnn.py:129,196


def scatter_mean(x_j: torch.Tensor, dst: torch.Tensor, n_dst: int):
    """PyG ``scatter(x_j, dst, 0, n_dst, reduce='mean')`` as used by ``MeanAggregation``.

    ``sum`` is a sequential ``scatter_add_`` in edge order on CPU; ``count.clamp(min=1)`` makes
    isolated destination rows exactly zero.  Returns ``(mean, count_clamped)``.
    """
    e, f = x_j.shape
    cnt = torch.zeros(n_dst, dtype=x_j.dtype).scatter_add_(0, dst, torch.ones(e, dtype=x_j.dtype))
    cnt = cnt.clamp(min=1)
    summ = torch.zeros(n_dst, f, dtype=x_j.dtype).scatter_add_(0, dst[:, None].expand(e, f), x_j)
    return summ / cnt[:, None], cnt


def sage_conv(x_src, x_dst, edge_index, w_l, b_l, w_r):
    """``SAGEConv((-1,-1), H)((x_src, x_dst), edge_index)`` -- train_gnn.py:177-184,194-197.

    flow = source_to_target: ``edge_index[0]`` indexes ``x_src``, ``edge_index[1]`` indexes
    ``x_dst``.  message(x_j) = x_j; aggregate = mean; update:
    ``lin_l(mean) + lin_r(x_dst)`` where ``lin_l`` has the bias and ``lin_r`` has none.
    E = 0 (inference.py:412-419) gives ``b_l + x_dst @ W_r^T``.
    """
    src, dst = edge_index[0], edge_index[1]
    x_j = x_src.index_select(0, src)
    mean, _ = scatter_mean(x_j, dst, x_dst.size(0))
    return F.linear(mean, w_l, b_l) + F.linear(x_dst, w_r)


class SAGEConvOracle(torch.nn.Module):
    """Module form with PyG's parameter names (``lin_l.weight``, ``lin_l.bias``, ``lin_r.weight``)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        mk = lambda c, bias: (torch.nn.LazyLinear(out_channels, bias=bias) if c < 0
                              else torch.nn.Linear(c, out_channels, bias=bias))
        self.lin_l = mk(in_channels[0], True)
        self.lin_r = mk(in_channels[1], False)

    def forward(self, x, edge_index):
        if isinstance(x, torch.Tensor):
            x = (x, x)
        x_src, x_dst = x
        src, dst = edge_index[0], edge_index[1]
        mean, _ = scatter_mean(x_src.index_select(0, src), dst, x_dst.size(0))
        return self.lin_l(mean) + self.lin_r(x_dst)


class WeightedRGCNOracle(torch.nn.Module):
    """train_gnn.py:147-200 verbatim (one hetero layer, fixed python-float weights).  ``conv_cls`` is the
    name the reference imports at train_gnn.py:6 (``SAGEConv``): the oracle's restatement by default; the
    drop-in tests bind the product's ``SAGEConv`` instead -- the import swap INTEGRATION.md describes --
    while every other line stays stock torch.  Pinned to the reference's own class text by
    tests/test_oracle.py (``ref_exec_*`` fixtures, tests/golden/make_golden_ref.py)."""

    def __init__(self, hidden_dim=64, in_channels=(-1, -1), conv_cls=None):
        super().__init__()
        conv_cls = conv_cls or SAGEConvOracle
        self.msg_direct = conv_cls(in_channels, hidden_dim)   # user <- post
        self.msg_social = conv_cls(in_channels, hidden_dim)   # user <- user
        self.post_update = conv_cls(in_channels, hidden_dim)  # post <- user
        self.w_direct = 1.0
        self.w_social = 0.75

    def forward(self, x_dict, edge_index_dict):
        user_x, post_x = x_dict["user"], x_dict["post"]
        msg_direct = self.msg_direct((post_x, user_x), edge_index_dict[REL_DIRECT])
        msg_social = self.msg_social((user_x, user_x), edge_index_dict[REL_SOCIAL])
        user_out = F.relu(self.w_direct * msg_direct + self.w_social * msg_social)
        post_out = F.relu(self.post_update((user_x, post_x), edge_index_dict[REL_ENGAGE]))
        return {"user": user_out, "post": post_out}


def _layers_of(model):
    return list(model.layers) if hasattr(model, "layers") else [model]


def pre_activations(layer, x_dict, edge_index_dict):
    """The two ReLU inputs of one ``WeightedRGCNOracle`` block (train_gnn.py:187-198 without the ReLU)."""
    user_x, post_x = x_dict["user"], x_dict["post"]
    z_u = (layer.w_direct * layer.msg_direct((post_x, user_x), edge_index_dict[REL_DIRECT])
           + layer.w_social * layer.msg_social((user_x, user_x), edge_index_dict[REL_SOCIAL]))
    z_p = layer.post_update((user_x, post_x), edge_index_dict[REL_ENGAGE])
    return z_u, z_p


def forward_gated(model, x_dict, edge_index_dict, gates=None):
    """Forward of ``model`` with the ReLU gates FORCED: ``gates[l] = {"user": bool[U,H], "post": bool[P,H]}``
    replaces ``relu(z)`` by ``z * gate`` in layer l (``None``: the ordinary ReLU).  ReLU is discontinuous
    in its derivative: a pre-activation within rounding of 0 can be gated either way by two correct fp32
    implementations, and ONE flipped gate moves a weight gradient by ~1e-3 of its scale.  Gradient parity is
    therefore stated as: (1) the gates of the code under test differ from the fp64 oracle's only where the
    fp64 pre-activation is within rounding of 0 (``z`` is returned for that check), and (2) given those
    gates, every gradient agrees to the stated tolerance.  Returns ``(out_dict, [(z_u, z_p) per layer])``."""
    zs = []
    for l, layer in enumerate(_layers_of(model)):
        z_u, z_p = pre_activations(layer, x_dict, edge_index_dict)
        zs.append((z_u, z_p))
        if gates is None:
            x_dict = {"user": F.relu(z_u), "post": F.relu(z_p)}
        else:
            x_dict = {"user": z_u * gates[l]["user"].to(z_u.dtype), "post": z_p * gates[l]["post"].to(z_p.dtype)}
    return x_dict, zs


def relu_margin(model, x_dict, edge_index_dict) -> float:
    """Smallest |pre-activation| / max|pre-activation| over every ReLU input of ``model`` (run it in fp64)."""
    margin = float("inf")
    with torch.no_grad():
        _, zs = forward_gated(model, x_dict, edge_index_dict)
    for pair in zs:
        for z in pair:
            if z.numel():
                margin = min(margin, float(z.abs().min() / z.abs().max().clamp(min=1e-30)))
    return margin


class StackedWeightedRGCNOracle(torch.nn.Module):
    """L stacked ``WeightedRGCN`` blocks (SURVEY.md §8 stacking rule; the reference has L = 1).

    Layer l is a fresh block with its own three convs fed the ``{'user','post'}`` output of
    layer l-1; ReLU after every layer including the last.
    """

    def __init__(self, hidden_dim=64, num_layers=2, in_channels=(-1, -1), conv_cls=None):
        super().__init__()
        self.layers = torch.nn.ModuleList(
            [WeightedRGCNOracle(hidden_dim, in_channels if i == 0 else (hidden_dim, hidden_dim), conv_cls)
             for i in range(num_layers)])

    def forward(self, x_dict, edge_index_dict):
        for layer in self.layers:
            x_dict = layer(x_dict, edge_index_dict)
        return x_dict


def link_loss(user_emb, post_emb, pos_u, pos_p, neg_p, interaction_type_tensor, num_users):
    """train_gnn.py:259-281 verbatim.  ``BCEWithLogitsLoss()`` has reduction='mean', so
    ``pos_loss`` is a SCALAR and ``(pos_weights * pos_loss).mean() == mean(w) * pos_loss``."""
    criterion = torch.nn.BCEWithLogitsLoss()
    pos_scores = (user_emb[pos_u] * post_emb[pos_p]).sum(dim=1)
    pos_global_post_ids = pos_p + num_users
    pos_weights = interaction_type_tensor[pos_global_post_ids]
    neg_scores = (user_emb[pos_u] * post_emb[neg_p]).sum(dim=1)
    pos_loss = criterion(pos_scores, torch.ones_like(pos_scores))
    neg_loss = criterion(neg_scores, torch.zeros_like(neg_scores))
    weighted_pos_loss = (pos_weights * pos_loss).mean()
    return weighted_pos_loss + neg_loss


def train_step(model, optimizer, x_dict, edge_index_dict, train_edge_index,
               interaction_type_tensor, num_users, num_posts, neg_p=None):
    """The body of ``train()`` train_gnn.py:242-285.  ``neg_p`` may be supplied so that both
    paths share the same negatives (the reference draws ``torch.randint(0, num_posts, (E,))``
    at :272)."""
    model.train()
    optimizer.zero_grad()
    out = model(x_dict, edge_index_dict)
    user_emb, post_emb = out["user"], out["post"]
    pos_u, pos_p = train_edge_index
    if neg_p is None:
        neg_p = torch.randint(0, num_posts, (pos_p.size(0),), device=pos_p.device)
    loss = link_loss(user_emb, post_emb, pos_u, pos_p, neg_p, interaction_type_tensor, num_users)
    loss.backward()
    optimizer.step()
    return loss.item()


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|, 1e-6*max|b|, tiny): relative error with an absolute floor, because
    post-ReLU outputs contain exact zeros (SURVEY.md §8c tolerances)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    scale = float(b.abs().max())
    floor = max(1e-6 * scale, 1e-30)
    return float(((a - b).abs() / b.abs().clamp(min=floor)).max())


def rel_err_norm(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max|b|: error relative to the tensor's scale (used for sums of many terms
    whose individual entries may cancel to ~0)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
