"""ctypes loader for oracle/graphbuilder_oracle.c and (when built) the reference's own Cython
builder in oracle/_ref -- TEST INFRASTRUCTURE ONLY (tests/, oracle/make_golden.py)."""
import ctypes as C
import importlib.util
import os
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _clib():
    path = os.path.join(HERE, "_build", "libgraphbuilder_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-C", HERE, "_build/libgraphbuilder_oracle.so"], check=True)
    lib = C.CDLL(path)
    lib.oracle_sliding_window.restype = C.c_uint32
    lib.oracle_sliding_window.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.oracle_edges_from_counts.restype = C.c_uint64
    lib.oracle_edges_from_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    return lib


def sliding_window(X, n_vocab, window_size):
    """graphbuilder.pyx:71-115 -> (packed c_ij uint32, n_windows)."""
    X = np.ascontiguousarray(X, dtype=np.int32)
    c = np.zeros(n_vocab * (n_vocab + 1) // 2, dtype=np.uint32)
    nw = _clib().oracle_sliding_window(X.ctypes.data, c.ctypes.data, window_size, n_vocab, X.shape[0], X.shape[1])
    return c, int(nw)


def compute_word_word_edges(X, n_vocab, window_size):
    """graphbuilder.pyx:23-66 restated: (int32[E,2], float32[E])."""
    c, nw = sliding_window(X, n_vocab, window_size)
    lib = _clib()
    n = lib.oracle_edges_from_counts(c.ctypes.data, n_vocab, nw, None, None)
    coo = np.empty((n, 2), dtype=np.int32)
    w = np.empty(n, dtype=np.float32)
    if n:
        lib.oracle_edges_from_counts(c.ctypes.data, n_vocab, nw, coo.ctypes.data, w.ctypes.data)
    return coo, w


def reference_module():
    """The reference's Cython graphbuilder compiled into oracle/_ref by oracle/Makefile, or None."""
    path = os.path.join(HERE, "_ref", "graphbuilder" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("graphbuilder", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
