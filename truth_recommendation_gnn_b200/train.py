"""The reference's training step and recommendation call, on the B200 kernels.

``train_step`` is the body of ``train()`` (train_gnn.py:242-285) with the same order of
operations: zero_grad -> forward -> positive/negative scoring -> loss -> backward -> Adam step.
``recommend`` is inference.py:427-428.
"""
from __future__ import annotations

import torch

from . import fused_step
from .functional import link_bce_loss, score_topk


_COPY_STREAMS: dict = {}


def stage_negatives(neg_p, device):
    """This step's host input (the sampled negatives, train_gnn.py:272, when they are drawn on the host):
    start the host -> device copy on a side stream so that it overlaps the forward pass, which does not
    need them.  Returns ``(device tensor, event)``; the consumer waits on the event right before the loss.
    Device tensors pass through (``event`` is None)."""
    if neg_p is None or neg_p.is_cuda or torch.device(device).type != "cuda":
        return neg_p, None       # (host features: the model itself raises TrgError -- there is no CPU path)
    device = torch.device(device)
    side = _COPY_STREAMS.get(device)
    if side is None:
        side = _COPY_STREAMS[device] = torch.cuda.Stream(device=device)
    with torch.cuda.stream(side):
        dev_t = neg_p.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
    dev_t.record_stream(torch.cuda.current_stream(device))
    return dev_t, ev


def train_step(model, optimizer, x_dict, edge_index_dict, train_edge_index, interaction_type_tensor,
               num_users, num_posts, neg_p=None, return_tensor=False, fused=None):
    """One full-batch step.  ``neg_p`` defaults to ``torch.randint(0, num_posts, (E,), device)`` as
    at train_gnn.py:272.  Returns ``loss.item()`` (train_gnn.py:285) or the 0-d device tensor when
    ``return_tensor`` (no host sync).

    ``fused``: ``None`` = take the tape-free forward+loss+backward (``fused_step.loss_and_grads``)
    whenever the model qualifies, ``False`` = always build the autograd tape and call
    ``loss.backward()`` (same kernels plus torch's element-wise ReLU-backward / accumulation passes),
    ``True`` = require the fused path."""
    model.train()
    optimizer.zero_grad()
    use_fused = fused_step.eligible(model, x_dict) if fused is None else bool(fused)
    neg_p, neg_ready = stage_negatives(neg_p, x_dict["user"].device)     # host negatives: copy overlaps the forward
    if use_fused:
        if neg_p is None:
            neg_p = torch.randint(0, num_posts, (train_edge_index.size(1),), device=x_dict["user"].device)
        loss = fused_step.loss_and_grads(model, x_dict, edge_index_dict, train_edge_index,
                                         interaction_type_tensor, num_users, neg_p, neg_ready=neg_ready)
    else:
        out = model(x_dict, edge_index_dict)
        user_emb, post_emb = out["user"], out["post"]
        if neg_ready is not None:
            torch.cuda.current_stream().wait_event(neg_ready)
        if neg_p is None:
            neg_p = torch.randint(0, num_posts, (train_edge_index.size(1),), device=user_emb.device)
        loss = link_bce_loss(user_emb, post_emb, train_edge_index, neg_p, interaction_type_tensor, num_users)
        loss.backward()
    optimizer.step()
    return loss.detach() if return_tensor else loss.item()


@torch.no_grad()
def recommend(user_emb, known_post_emb, k=10, id_offset=0):
    """``scores = torch.mm(user_emb, known_post_emb.T); torch.topk(scores, min(K, len(scores)))``
    (inference.py:427-428) for one user row or a batch.  Returns ``(top scores, top post ids)``."""
    return score_topk(user_emb, known_post_emb, k, id_offset)
