"""Seeded synthetic user-post graphs in the reference's own formats.

The reference's dataset (``truth_social/*.tsv``, build_graph.py:107,160,185,277) is not
available, so every test and bench uses graphs drawn here.  What is mirrored:

* node features: unit-norm fp32 rows, as produced by the random projection + row L2
  normalisation at build_graph.py:452-456;
* ``edge_index[2, E]`` int64 COO with LOCAL ids per node type, unsorted, duplicates kept
  (build_graph.py:387,394,402; train_gnn.py:128-133); ``rev_engages = engages.flip(0)``
  (train_gnn.py:142);
* the interaction-type weight table indexed by GLOBAL post id ``num_users + local``
  (train_gnn.py:226-237): 3.0 for quotes ("QT"), 1.0 otherwise, 0.0 for ids never seen.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

REL_DIRECT = ("post", "rev_engages", "user")
REL_SOCIAL = ("user", "social", "user")
REL_ENGAGE = ("user", "engages", "post")

#: BASELINE.json configs -> (U, P, E_eng, E_soc, F=H, L).  E split 0.8/0.2.
CONFIGS = {
    "cfg1": dict(num_users=10_000, num_posts=50_000, e_eng=400_000, e_soc=100_000, hidden=64, layers=2),
    "cfg2": dict(num_users=1_000_000, num_posts=5_000_000, e_eng=40_000_000, e_soc=10_000_000, hidden=128, layers=2),
    "cfg4": dict(num_users=10_000_000, num_posts=50_000_000, e_eng=800_000_000, e_soc=200_000_000, hidden=256, layers=3),
}


@dataclass
class SynthGraph:
    x_dict: dict
    edge_index_dict: dict
    train_edge_index: torch.Tensor          # == edge_index_dict[REL_ENGAGE] (train_gnn.py:213)
    interaction_type_tensor: torch.Tensor   # [U + P] fp32
    num_users: int
    num_posts: int

    def to(self, device, non_blocking=False):
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        ei = {k: mv(v) for k, v in self.edge_index_dict.items()}
        return SynthGraph({k: mv(v) for k, v in self.x_dict.items()}, ei, ei[REL_ENGAGE],
                          mv(self.interaction_type_tensor), self.num_users, self.num_posts)

    @property
    def mp_edges(self) -> int:
        """Message-passing edges per hetero layer: 2*E_eng + E_soc."""
        return sum(int(v.size(1)) for v in self.edge_index_dict.values())


def _randint(hi, n, g, skew, device):
    if not skew:
        return torch.randint(0, max(hi, 1), (n,), generator=g, device=device)
    # Zipf-like: floor(N * u^3), u ~ U(0,1): a few destinations collect most edges
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    return (hi * u.pow(3)).long().clamp_(0, max(hi - 1, 0))


def synth_graph(num_users, num_posts, e_eng, e_soc, feat, seed=0, skew=False, device="cpu",
                dtype=torch.float32) -> SynthGraph:
    """Draw order is fixed: user feats, post feats, engage src/dst, social src/dst, weights."""
    g = torch.Generator(device=device).manual_seed(seed)
    xu = F.normalize(torch.randn(num_users, feat, generator=g, device=device), dim=1)
    xp = F.normalize(torch.randn(num_posts, feat, generator=g, device=device), dim=1)
    eng = torch.stack([_randint(num_users, e_eng, g, False, device),
                       _randint(num_posts, e_eng, g, skew, device)])
    soc = torch.stack([_randint(num_users, e_soc, g, False, device),
                       _randint(num_users, e_soc, g, skew, device)])
    gw = torch.Generator(device=device).manual_seed(seed + 2)
    w = torch.zeros(num_users + num_posts, device=device)
    w[num_users:] = torch.where(torch.rand(num_posts, generator=gw, device=device) < 0.25, 3.0, 1.0)
    ei = {REL_SOCIAL: soc, REL_ENGAGE: eng, REL_DIRECT: eng.flip(0).contiguous()}
    return SynthGraph({"user": xu.to(dtype), "post": xp.to(dtype)}, ei, eng, w, num_users, num_posts)


def synth_neg(num_posts, e_eng, step, device="cpu"):
    """``torch.randint(0, num_posts, (E,))`` of train_gnn.py:272 with a per-step seed so both
    paths share the negatives."""
    g = torch.Generator(device=device).manual_seed(3 + step)
    return torch.randint(0, num_posts, (e_eng,), generator=g, device=device)


def synth_queries(batch, num_posts, hidden, seed=4, zero_frac=0.0, device="cpu", dtype=torch.float32):
    """Config-5 inputs: post-ReLU (non-negative) query and catalogue rows; ``zero_frac`` of the
    catalogue rows are all-zero (dead ReLU) to stress tie handling."""
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.relu(torch.randn(batch, hidden, generator=g, device=device))
    cat = torch.relu(torch.randn(num_posts, hidden, generator=g, device=device))
    if zero_frac > 0:
        dead = torch.rand(num_posts, generator=g, device=device) < zero_frac
        cat[dead] = 0
    return q.to(dtype), cat.to(dtype)


def init_state_dict(hidden, feat, num_layers=1, seed=1):
    """``torch.nn.Linear``-default init (what PyG's ``Linear`` uses for SAGEConv) drawn in the
    order direct, social, post_update, layer by layer.  Keys follow the reference checkpoint
    (train_gnn.py:410): ``<conv>.lin_l.weight/bias``, ``<conv>.lin_r.weight``; stacked models
    prefix ``layers.<l>.``."""
    torch.manual_seed(seed)
    sd = {}
    for l in range(num_layers):
        fin = feat if l == 0 else hidden
        pre = f"layers.{l}." if num_layers > 1 else ""
        for conv in ("msg_direct", "msg_social", "post_update"):
            lin_l = torch.nn.Linear(fin, hidden, bias=True)
            lin_r = torch.nn.Linear(fin, hidden, bias=False)
            sd[f"{pre}{conv}.lin_l.weight"] = lin_l.weight.detach().clone()
            sd[f"{pre}{conv}.lin_l.bias"] = lin_l.bias.detach().clone()
            sd[f"{pre}{conv}.lin_r.weight"] = lin_r.weight.detach().clone()
    return sd


# ------------------------------------------------------------------------------------------------
# counter-based graphs: every edge / feature row is a pure function of its index
# ------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    """Python int -> the int64 with the same low 64 bits (torch int64 arithmetic wraps)."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """Logical right shift of int64 (``>>`` on a signed tensor is arithmetic)."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def hash64(idx: torch.Tensor, salt: int) -> torch.Tensor:
    """splitmix64 finaliser of ``idx + salt * golden``: a counter-based generator -- element i of a stream
    depends on i alone, so any rank can produce any slice of it without producing the rest."""
    z = idx + _i64(0x9E3779B97F4A7C15 * (2 * salt + 1))
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


class CounterGraph:
    """A synthetic user-post graph of the SynthGraph family whose parts can be generated independently:
    edge e of a relation is ``(hash(e, a) mod N_src, hash(e, b) mod N_dst)`` (uniform endpoints, duplicates
    kept), feature rows come in fixed blocks of ``ROW_BLOCK`` rows each seeded by its block id, and the
    interaction weight of a post is a hash of its id.  BASELINE config 4 (1B edges, H = 256: 30 GB of
    features and 16 GB of COO) cannot round-trip through one host or one GPU; with this
    generator every rank produces exactly the rows and edges it owns, on its own device
    (``dist.ShardedGraph.from_generator``).  ``materialize`` builds the whole graph for sizes where that is
    possible -- the tests compare the sharded-from-generator path with it."""

    ROW_BLOCK = 1 << 16

    def __init__(self, num_users, num_posts, e_eng, e_soc, feat, seed=0):
        self.num_users, self.num_posts = int(num_users), int(num_posts)
        self.e_eng, self.e_soc, self.feat, self.seed = int(e_eng), int(e_soc), int(feat), int(seed)

    def n_edges(self, rel):
        return self.e_soc if rel == REL_SOCIAL else self.e_eng

    def edges(self, rel, a, b, device):
        """``edge_index[:, a:b]`` of relation ``rel`` (int64 ``[2, b - a]``)."""
        e = torch.arange(a, b, device=device, dtype=torch.int64)
        s0 = 16 * self.seed
        if rel == REL_SOCIAL:
            return torch.stack([_lsr(hash64(e, s0 + 3), 1) % self.num_users,
                                _lsr(hash64(e, s0 + 4), 1) % self.num_users])
        u = _lsr(hash64(e, s0 + 1), 1) % self.num_users
        p = _lsr(hash64(e, s0 + 2), 1) % self.num_posts
        return torch.stack([u, p]) if rel == REL_ENGAGE else torch.stack([p, u])

    def features(self, node_type, r0, r1, device, dtype=torch.float32):
        """Unit-norm rows ``[r0, r1)`` of a node type's feature table (build_graph.py:452-456 normalises rows)."""
        if r1 <= r0:
            return torch.empty(0, self.feat, device=device, dtype=dtype)
        rb = self.ROW_BLOCK
        out = []
        for blk in range(r0 // rb, (r1 - 1) // rb + 1):
            g = torch.Generator(device=device).manual_seed(
                (self.seed * 1_000_003 + (0 if node_type == "user" else 500_000_000) + blk) & 0x7FFFFFFFFFFFFFFF)
            x = F.normalize(torch.randn(rb, self.feat, generator=g, device=device), dim=1)
            lo, hi = max(r0, blk * rb) - blk * rb, min(r1, (blk + 1) * rb) - blk * rb
            out.append(x[lo:hi].to(dtype))
        return torch.cat(out) if len(out) > 1 else out[0]

    def post_weight(self, post_ids):
        """interaction-type weight of posts (train_gnn.py:226-237): 3.0 ("QT") with probability 1/4, else 1.0."""
        h = _lsr(hash64(post_ids, 16 * self.seed + 5), 1) % 4
        return torch.where(h == 0, 3.0, 1.0).float()

    def materialize(self, device="cpu", dtype=torch.float32) -> SynthGraph:
        eng = self.edges(REL_ENGAGE, 0, self.e_eng, device)
        soc = self.edges(REL_SOCIAL, 0, self.e_soc, device)
        w = torch.zeros(self.num_users + self.num_posts, device=device)
        w[self.num_users:] = self.post_weight(torch.arange(self.num_posts, device=device))
        ei = {REL_SOCIAL: soc, REL_ENGAGE: eng, REL_DIRECT: eng.flip(0).contiguous()}
        x = {"user": self.features("user", 0, self.num_users, device, dtype),
             "post": self.features("post", 0, self.num_posts, device, dtype)}
        return SynthGraph(x, ei, eng, w, self.num_users, self.num_posts)
