"""ctypes binding of the C ABI declared in include/textgcn_b200.h.

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised (the reference's sweeps catch RuntimeError, old/h_o_train.py:129-131).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_LIB_NAME = "libtextgcn_b200.so"
_LIB_PATH = os.environ.get("TGCN_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", _LIB_NAME)
_lib: Optional[C.CDLL] = None

c_i32p = C.POINTER(C.c_int32)
c_void = C.c_void_p


class SpmmArgs(C.Structure):
    """Mirror of tgcn_spmm_args (include/textgcn_b200.h)."""
    _fields_ = [
        ("rowptr", c_void), ("colidx", c_void), ("val", c_void),
        ("chunks", c_void), ("n_chunks", C.c_int32),
        ("split_rows", c_void), ("n_split_rows", C.c_int32),
        ("slot_owner", c_void), ("split_counters", c_void),
        ("scratch", c_void),
        ("B", c_void), ("ldb", C.c_int64), ("b_dtype", C.c_int32),
        ("C", c_void), ("ldc", C.c_int64), ("c_dtype", C.c_int32),
        ("F", C.c_int32),
        ("c_row_offset", C.c_int64),
        ("bias", c_void), ("bias_len", C.c_int32),
        ("act", C.c_int32),
        ("drop_mode", C.c_int32), ("drop_p", C.c_float), ("keep_mask", c_void), ("ldmask", C.c_int64),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64), ("philox_offset_dev", c_void), ("philox_row_offset", C.c_int64),
        ("W_proj", c_void), ("n_proj", C.c_int32), ("P", c_void), ("ldp", C.c_int64),
        ("adam_param", c_void), ("adam_exp_avg", c_void), ("adam_exp_avg_sq", c_void), ("adam_max_exp_avg_sq", c_void),
        ("adam_ld", C.c_int64), ("adam_hyper_dev", c_void), ("adam_beta1", C.c_float), ("adam_beta2", C.c_float),
        ("adam_eps", C.c_float), ("adam_param_mirror_mc", c_void),
        ("tc_part", c_void), ("tc_ld", C.c_int64), ("tc_rank", c_void), ("tc_slot_ptr", c_void),
        ("raw_in", c_void), ("raw_ld", C.c_int64), ("raw_stride", C.c_int64), ("n_raw", C.c_int32), ("raw_rows", C.c_int64),
        ("adam_mirror_rows", C.c_int64), ("c_scatter_bases", c_void), ("c_scatter_rows", C.c_int64), ("c_scatter_row0", C.c_int64),
    ]


class TcPlanArgs(C.Structure):
    """Mirror of tgcn_tc_plan."""
    _fields_ = [
        ("A_tiles", c_void), ("tile_kb", c_void), ("n_tiles", C.c_int32), ("units", c_void), ("n_units", C.c_int32),
        ("perm", c_void), ("n_col_blocks", C.c_int32),
    ]


class ProjectArgs(C.Structure):
    """Mirror of tgcn_project_args."""
    _fields_ = [
        ("X", c_void), ("ldx", C.c_int64), ("x_dtype", C.c_int32), ("n_rows", C.c_int64), ("K", C.c_int32),
        ("W", c_void), ("M", C.c_int32), ("bias", c_void),
        ("P", c_void), ("ldp", C.c_int64), ("P_mirror_mc", c_void),
        ("drop_mode", C.c_int32), ("drop_p", C.c_float), ("keep_mask", c_void), ("ldmask", C.c_int64),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64), ("philox_offset_dev", c_void), ("philox_row_offset", C.c_int64),
        ("Xd", c_void), ("ldxd", C.c_int64), ("mirror_rows", C.c_int64),
    ]


class DenseBwdArgs(C.Structure):
    """Mirror of tgcn_dense_bwd_args."""
    _fields_ = [
        ("G2", c_void), ("ldg2", C.c_int64),
        ("H1d", c_void), ("ldh", C.c_int64), ("h_dtype", C.c_int32),
        ("W2", c_void),
        ("dZ2", c_void), ("lddz2", C.c_int64),
        ("n_rows", C.c_int64), ("row_offset", C.c_int64),
        ("H", C.c_int32), ("C", C.c_int32),
        ("act", C.c_int32), ("drop_mode", C.c_int32), ("drop_p", C.c_float),
        ("keep_mask", c_void), ("ldmask", C.c_int64),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64), ("philox_offset_dev", c_void),
        ("dZ1", c_void), ("lddz1", C.c_int64), ("dz1_dtype", C.c_int32),
        ("dZ1_mirror_mc", c_void),
        ("dW2", c_void), ("db_hidden", c_void), ("db_out", c_void), ("dZ1_mirror_rows", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/textgcn_b200.h declares
SIGNATURES = {
    "tgcn_last_error": (C.c_char_p, []),
    "tgcn_version": (C.c_int, []),
    "tgcn_launch_count": (C.c_uint64, []),
    "tgcn_device_info": (C.c_int, [c_i32p, c_i32p, c_i32p]),
    "tgcn_csr_workspace_bytes": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(C.c_size_t)]),
    "tgcn_csr_from_coo_gcn_norm": (C.c_int, [c_void, c_void, C.c_int64, c_void, C.c_int64, C.c_int64,
                                             c_void, c_void, c_void, c_void, c_void, c_void,
                                             c_void, C.c_size_t, c_void]),
    "tgcn_spmm_plan": (C.c_int, [c_void, C.c_int64, C.c_int64, C.c_int32, c_void, C.c_int64, c_void, c_void, c_void,
                                 c_void, C.c_size_t, c_void]),
    "tgcn_spmm_plan_workspace_bytes": (C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    "tgcn_spmm": (C.c_int, [C.POINTER(SpmmArgs), c_void]),
    "tgcn_spmm_tc": (C.c_int, [C.POINTER(TcPlanArgs), c_void, C.c_int64, C.c_int32, c_void, c_void, C.c_int64, c_void]),
    "tgcn_spmm_tc_workspace_elems": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(C.c_int64)]),
    "tgcn_masked_nll": (C.c_int, [c_void, C.c_int64, C.c_int64, C.c_int32, c_void, c_void, C.c_int64,
                                  c_void, c_void, c_void, C.c_int64, c_void, c_void, c_void, c_void, c_void,
                                  c_void, C.c_size_t, c_void]),
    "tgcn_masked_nll_workspace_bytes": (C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    "tgcn_dense_bwd": (C.c_int, [C.POINTER(DenseBwdArgs), c_void, C.c_size_t, c_void]),
    "tgcn_dense_bwd_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "tgcn_project": (C.c_int, [c_void, C.c_int64, C.c_int32, C.c_int64, C.c_int32, c_void, C.c_int32, c_void,
                               c_void, C.c_int64, c_void, c_void]),
    "tgcn_project_ex": (C.c_int, [C.POINTER(ProjectArgs), c_void]),
    "tgcn_colsum": (C.c_int, [c_void, C.c_int64, C.c_int64, C.c_int32, c_void, c_void, C.c_size_t, c_void]),
    "tgcn_colsum_workspace_bytes": (C.c_int, [C.c_int32, C.POINTER(C.c_size_t)]),
    "tgcn_dropout_apply": (C.c_int, [c_void, C.c_int64, c_void, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                     c_void, C.c_int64, C.c_uint64, C.c_uint64, c_void, C.c_int64, c_void]),
    "tgcn_hier_forward": (C.c_int, [c_void, C.c_int64, C.c_int64, C.c_int64, c_void, C.c_int64, C.c_int32,
                                    C.c_int32, c_void, C.c_int64, c_void]),
    "tgcn_hier_backward": (C.c_int, [c_void, C.c_int64, C.c_int64, C.c_int64, c_void, C.c_int64, C.c_int32,
                                     C.c_int32, c_void, c_void, C.c_size_t, c_void]),
    "tgcn_hier_backward_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "tgcn_adam_step": (C.c_int, [c_void, c_void, c_void, c_void, c_void, C.c_int64, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_int32, C.c_int64, c_void, c_void, c_void]),
    "tgcn_increment_step": (C.c_int, [c_void, c_void]),
    "tgcn_adam_prepare": (C.c_int, [c_void, c_void, C.c_float, C.c_float, C.c_float, c_void]),
    "tgcn_adam_step_small": (C.c_int, [C.c_int32, c_void, c_void, c_void, c_void, c_void, c_void, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_int32, C.c_int64, c_void, c_void]),
    "tgcn_peer_push": (C.c_int, [c_void, c_void, C.c_int32, C.c_int32, C.c_int64, C.c_int64, c_void, c_void]),
    "tgcn_sum_slots": (C.c_int, [c_void, C.c_int32, C.c_int64, C.c_int64, c_void, c_void]),
    "tgcn_count_mask": (C.c_int, [c_void, C.c_int64, c_void, c_void]),
    "tgcn_cast_f32_to_bf16": (C.c_int, [c_void, c_void, C.c_int64, c_void]),
}


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_NAME} not found at {_LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pytextgcn_b200/csrc`.  pytextgcn_b200 has no CPU / eager fallback.")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError -> missing export, surfaced to the caller
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().tgcn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"textgcn_b200 error {rc}: {msg}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
