"""ctypes binding of ``oracle/csrc/oracle_int.c`` (built into ``oracle/_ref/``).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle_int.so")


def build():
    src = os.path.join(_HERE, "csrc", "oracle_int.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def csr_by_dst(edge_index: torch.Tensor, n_dst: int):
    src = edge_index[0].contiguous().long()
    dst = edge_index[1].contiguous().long()
    e = src.numel()
    rowptr = torch.empty(n_dst + 1, dtype=torch.long)
    col = torch.empty(e, dtype=torch.long)
    eid = torch.empty(e, dtype=torch.long)
    rc = lib().oracle_csr_by_dst(_p(src), _p(dst), ctypes.c_int64(e), ctypes.c_int64(n_dst),
                                 _p(rowptr), _p(col), _p(eid))
    if rc:
        raise ValueError(f"oracle_csr_by_dst rc={rc}")
    return rowptr, col, eid


def topk_rows(scores: torch.Tensor, k: int, id_offset: int = 0):
    scores = scores.contiguous().float()
    b, n = scores.shape
    k = min(k, n)
    vals = torch.empty(b, k)
    ids = torch.empty(b, k, dtype=torch.long)
    for r in range(b):
        lib().oracle_topk_row(_p(scores[r]), ctypes.c_int64(n), ctypes.c_int64(k),
                              ctypes.c_int64(id_offset), _p(vals[r]), _p(ids[r]))
    return vals, ids
