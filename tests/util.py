"""Shared helpers for the parity tests (tolerances are BASELINE.json's north_star)."""
import os

import torch

from oracle import sage as osage
from truth_recommendation_gnn_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_F32 = 1e-5   # relative, fp32 embeddings / losses / scores
TOL_BF16 = 1e-2  # relative, bf16 storage with fp32 accumulation


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def golden_graph(fix, device="cpu"):
    ei = {eval(k): v.to(device) for k, v in fix["edge_index"].items()}
    m = fix["meta"]
    return synth.SynthGraph({"user": fix["x_user"].to(device), "post": fix["x_post"].to(device)}, ei,
                            ei[synth.REL_ENGAGE], fix["w"].to(device), m["u"], m["p"])


def oracle_model(h, layers, sd, fin=None):
    fin = fin or h
    m = (osage.WeightedRGCNOracle(h, (fin, fin)) if layers == 1
         else osage.StackedWeightedRGCNOracle(h, layers, (fin, fin)))
    m.load_state_dict(sd)
    return m


def assert_close(a, b, tol, what=""):
    """Error relative to the tensor's scale (max|b|): sums of many signed terms may cancel to ~0
    elementwise, so the elementwise-relative form is only used where entries are O(1)."""
    err = osage.rel_err_norm(a, b)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol:.1e}"
    return err
