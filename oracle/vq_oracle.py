"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/vq_oracle.c (CPU VQ/LBG oracle).

Never imported by the product path (hmm_training_b200/*); see oracle/vq_oracle.c header.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvq_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                               "-Wall", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        lib.vqo_encode.restype = ctypes.c_int
        lib.vqo_encode.argtypes = [dp, ctypes.c_long, dp, ctypes.c_int, ip, dp]
        lib.vqo_lbg.restype = ctypes.c_int
        lib.vqo_lbg.argtypes = [dp, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                dp, dp, ip, ip, dp]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def encode(X: np.ndarray, C: np.ndarray, return_dist: bool = False):
    """HMM/hmm_training.py:95-118 on a flat [F,13] matrix -> int32 indices [F]."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    C = np.ascontiguousarray(C, dtype=np.float64)
    assert X.ndim == 2 and X.shape[1] == 13 and C.shape[1] == 13
    idx = np.empty(X.shape[0], dtype=np.int32)
    dist = np.empty(X.shape[0], dtype=np.float64)
    rc = _load().vqo_encode(_dp(X), X.shape[0], _dp(C), C.shape[0], _ip(idx), _dp(dist))
    if rc != 0:
        raise RuntimeError(f"vqo_encode failed: {rc}")
    return (idx, dist) if return_dist else idx


def lbg(X: np.ndarray, K: int = 256, max_iterations: int = 100, epsilon: float = 0.001):
    """CodeVector/codevector_functions.py:442-531 (createCodeVector).  Returns
    (centroids [Kout,13], generations list of arrays, assign [F], iters_per_gen, gdist)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.shape[0] == 0:
        raise ValueError("No raw data provided")
    n_gen = int(np.log2(K))
    kout = 1 << max(n_gen, 1)
    n_rows = 1 + sum(1 << g for g in range(1, n_gen + 1))
    C = np.zeros((kout, 13))
    gens = np.zeros((n_rows, 13))
    assign = np.zeros(X.shape[0], dtype=np.int32)
    iters = np.zeros(max(n_gen, 1), dtype=np.int32)
    gdist = np.zeros(max(n_gen, 1))
    rc = _load().vqo_lbg(_dp(X), X.shape[0], K, max_iterations, epsilon, _dp(C), _dp(gens),
                         _ip(assign), _ip(iters), _dp(gdist))
    if rc < 0:
        raise RuntimeError(f"vqo_lbg failed: {rc}")
    out, pos = [gens[0:1].copy()], 1
    for g in range(1, n_gen + 1):
        out.append(gens[pos:pos + (1 << g)].copy())
        pos += 1 << g
    return C[:rc].copy(), out, assign, iters[:n_gen], gdist[:n_gen]
