"""Python face of the native word-word PMI edge builder (pytextgcn_b200/csrc_host/graph_builder.cpp).

Same call signature and return types as the reference's Cython entry point
`compute_word_word_edges(X, n_vocab, n_documents, seq_len, window_size=20, n_jobs=1, verbose=0)`
(textgcn/lib/clib/graphbuilder.pyx:23-66): returns (int32[E, 2] COO, float32[E] PMI weights), edges
emitted as (i,j),(j,i) pairs in upper-triangle row-major order.  Differences: `n_jobs` is honoured
(threads over documents; the reference documents it as UNUSED, graphbuilder.pyx:36), memory is
O(#co-occurring pairs) instead of O(V^2), no V < 65,536 limit, and the returned arrays own their
memory (the reference's alias malloc'd buffers that are never freed, graphbuilder.pyx:65-66).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import numpy as np

_LIB = None
_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtextgcn_host.so")


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} not found: build it with `make -C pytextgcn_b200/csrc_host` "
                               "(or __graft_entry__.build())")
        lib = C.CDLL(_PATH)
        lib.tgcn_ww_build.restype = C.c_void_p
        lib.tgcn_ww_build.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
        lib.tgcn_ww_fetch.restype = C.c_int
        lib.tgcn_ww_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.tgcn_ww_free.restype = None
        lib.tgcn_ww_free.argtypes = [C.c_void_p]
        lib.tgcn_ww_counts_packed.restype = C.c_int
        lib.tgcn_ww_counts_packed.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                              C.POINTER(C.c_uint64)]
        _LIB = lib
    return _LIB


def _check_tokens(X: np.ndarray, n_documents: int, seq_len: int) -> np.ndarray:
    X = np.ascontiguousarray(X, dtype=np.int32)
    if X.ndim != 2 or X.shape != (n_documents, seq_len):
        raise ValueError(f"X must be int32 of shape (n_documents, seq_len) = ({n_documents}, {seq_len}), got {X.shape}")
    return X


def compute_word_word_edges(X, n_vocab: int, n_documents: int, seq_len: int, window_size: int = 20,
                            n_jobs: int = 1, verbose: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    X = _check_tokens(X, n_documents, seq_len)
    lib = _lib()
    n_edges, n_win = C.c_int64(0), C.c_uint64(0)
    h = lib.tgcn_ww_build(X.ctypes.data, n_documents, seq_len, n_vocab, window_size, int(n_jobs),
                          C.byref(n_edges), C.byref(n_win))
    if not h:
        raise RuntimeError("compute_word_word_edges: bad input (token id outside [0, n_vocab) or empty shape)")
    try:
        coo = np.empty((n_edges.value, 2), dtype=np.int32)
        w = np.empty(n_edges.value, dtype=np.float32)
        if n_edges.value and lib.tgcn_ww_fetch(h, coo.ctypes.data, w.ctypes.data) != 0:
            raise RuntimeError("compute_word_word_edges: fetch failed")
    finally:
        lib.tgcn_ww_free(h)
    if verbose > 1:
        print(f"Number of word-word-edges: {n_edges.value} ({n_win.value} windows)")
    return coo, w


def sliding_window_tester(X, n_vocab: int, n_documents: int, seq_len: int, window_size: int = 20,
                          n_jobs: int = 1) -> np.ndarray:
    """Packed upper-triangular pair counts, like the reference's test shim (graphbuilder.pyx:263-275)."""
    X = _check_tokens(X, n_documents, seq_len)
    out = np.zeros(n_vocab * (n_vocab + 1) // 2, dtype=np.uint32)
    n_win = C.c_uint64(0)
    rc = _lib().tgcn_ww_counts_packed(X.ctypes.data, n_documents, seq_len, n_vocab, window_size, out.ctypes.data,
                                      C.byref(n_win))
    if rc != 0:
        raise RuntimeError("sliding_window_tester: bad input")
    return out
