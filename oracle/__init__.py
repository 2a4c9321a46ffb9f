"""CPU oracle for the hetero-SAGE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the CPU
baseline -- never as something the product path calls.  The product package
(``truth_recommendation_gnn_b200``) does not import it and raises when its CUDA library is
missing; there is no CPU fallback.

What it restates (citations are into ``/root/reference``):

* ``sage.sage_conv`` / ``sage.SAGEConvOracle`` -- ``torch_geometric.nn.SAGEConv`` with the
  options the reference uses (``train_gnn.py:158-160``: ``aggr='mean'``, ``root_weight=True``,
  ``bias=True``, ``normalize=False``, ``project=False``), called at ``train_gnn.py:177-184,194-197``.
* ``sage.WeightedRGCNOracle`` -- ``train_gnn.py:147-200`` verbatim (dup ``inference.py:119-169``).
* ``sage.train_step`` -- the body of ``train()`` ``train_gnn.py:242-285`` verbatim, including
  the scalar-loss quirk at ``:276-281``.
* ``topk.score_topk`` -- ``inference.py:427-428`` (``torch.mm`` + ``torch.topk``) with the
  canonical tie rule (score desc, id asc).
* ``csr.csr_by_dst`` -- the bit-exact contract for the destination-sorted CSR built from the
  COO ``edge_index`` of ``build_graph.py:387,394,402`` / ``train_gnn.py:128-142``.
* ``csrc/oracle_int.c`` -- plain-C restatement of the two integer algorithms (stable counting
  sort CSR, canonical top-k) used to cross-check the torch versions.

PARITY PIN STATUS
-----------------
* torch parts (``mm``, ``topk`` values, ``BCEWithLogitsLoss``, ``Adam``, ``index``/``scatter_add_``)
  are executed by the very same torch the reference would call: pinned.
* ``SAGEConv`` lives in the third-party ``torch_geometric`` (PyPI ``torch-geometric``; the
  reference pins NO version -- no requirements/lock file exists) which is absent from
  ``/root/reference``, not installed here and not installable (no network).  The reference has
  no tests, golden vectors or fixtures.  For that operator this oracle is therefore
  **"parity unpinned"**: it restates PyG 2.x's published algorithm
  (``MessagePassing.propagate`` -> ``index_select`` + ``scatter(reduce='mean')`` with
  ``count.clamp(min=1)``; ``out = lin_l(mean) + lin_r(x_dst)``) and is anchored on the
  reference's call sites and on dense-adjacency math (``tests/test_oracle.py``), plus the
  committed fixtures under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
* everything AROUND that operator is pinned to the reference's own text:
  ``tests/golden/make_golden_ref.py`` extracts ``WeightedRGCN``, ``train()``, ``evaluate()`` and
  ``build_edge_index_safe`` from ``/root/reference/train_gnn.py`` by name (``ast``) and executes them
  unmodified with ``SAGEConv`` bound to ``sage.SAGEConvOracle``; the outputs are the committed
  ``tests/golden/ref_exec_*.pt`` fixtures that ``tests/test_oracle.py`` holds this package to.
"""
from . import csr, sage, topk  # noqa: F401  (oracle.evaluate needs sklearn: imported on demand)
