"""Thin Python wrappers over the C ABI (one function per entry point of
include/textgcn_b200.h).  All tensors must be CUDA tensors; nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _native
from .graph import GraphCSR, SpmmPlan

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU = 0, 1
DROP_NONE, DROP_MASK, DROP_PHILOX = 0, 1, 2
# Layer 2's thin projection can run inside the layer-1 SpMM epilogue (tgcn_spmm W_proj/P, register
# butterfly) or as the stand-alone 4-row register-blocked kernel (tgcn_project).  Measured on B200
# (ms/epoch, fused vs stand-alone): 20NG-shape 3.97 vs 3.83, Amazon-shape (64 classes) 2.67 vs 1.93,
# R8 0.704 vs 0.699 -- the epilogue's per-lane weight-row reads are bank-conflicted and serialise the
# tail of every row warp, while the stand-alone kernel re-reads H1d (L2-resident) for ~40 us.  So the
# shipped default is stand-alone; classes <= this threshold use the fused epilogue (0 = never).
import os as _os
FUSED_PROJ_MAX_CLASSES = int(_os.environ.get("TGCN_FUSED_PROJ_MAX_CLASSES", "0"))
_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {t.dtype} (fp32 or bf16 only)")


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pytextgcn_b200 ops need CUDA tensors (there is no CPU path)")


def pad4(n: int) -> int:
    return (n + 3) & ~3


def pad8(n: int) -> int:
    return (n + 7) & ~7


def spmm(graph: GraphCSR, B: torch.Tensor, *, F: Optional[int] = None, out: Optional[torch.Tensor] = None,
         out_dtype: Optional[torch.dtype] = None, plan: Optional[SpmmPlan] = None,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         drop_mode: int = DROP_NONE, drop_p: float = 0.0, keep_mask: Optional[torch.Tensor] = None,
         philox_seed: int = 0, philox_offset: int = 0, philox_offset_dev: Optional[torch.Tensor] = None, row_id_offset: int = 0,
         W_proj: Optional[torch.Tensor] = None, P: Optional[torch.Tensor] = None,
         want_out: bool = True, adam: Optional[dict] = None, tc=None,
         raw_slots: Optional[torch.Tensor] = None, scatter: Optional[dict] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """C = epi(A_hat[rows of plan] @ B[:, :F]) (+ P = C @ W_proj).  See tgcn_spmm.
    tc: a TcPlan whose dense-tile partial rows (already computed by tgcn_spmm_tc for this B, see spmm_hybrid) are
    added to every row before the epilogue; `graph` must then be the plan's remainder.

    B: [>= n_nodes, >= F] fp32/bf16, row stride a multiple of 4 (fp32) / 8 (bf16) elements.
    Returns (C or None, P or None).  C has plan.row_end - plan.row_begin rows.
    """
    _need_cuda(B, out, bias, keep_mask, W_proj, P)
    lib = _native.load()
    plan = plan or graph.plan()
    F = int(B.shape[1]) if F is None else int(F)
    n_out = plan.row_end - plan.row_begin
    if B.dim() != 2 or B.stride(1) != 1:
        raise RuntimeError("spmm: B must be a 2-D row-major tensor")
    if B.shape[0] < graph.n_cols:
        raise RuntimeError(f"spmm: B has {B.shape[0]} rows, the graph has {graph.n_cols} columns")
    a = _native.SpmmArgs()
    if graph.nnz == 0:      # e.g. the remainder of a hybrid plan whose every entry sits in a dense tile: rows are all empty
        if "_empty_entries" not in graph.buffers:
            graph.buffers["_empty_entries"] = torch.zeros(4, dtype=torch.int32, device=graph.device)
        e = graph.buffers["_empty_entries"]
        a.rowptr, a.colidx, a.val = graph.rowptr.data_ptr(), e.data_ptr(), e.data_ptr()
    else:
        a.rowptr, a.colidx, a.val = graph.rowptr.data_ptr(), graph.colidx.data_ptr(), graph.val.data_ptr()
    a.chunks, a.n_chunks = plan.chunks.data_ptr(), plan.n_chunks
    a.split_rows, a.n_split_rows = (plan.split_rows.data_ptr() if plan.n_split_rows else None), plan.n_split_rows
    scratch = plan.scratch(F)
    a.scratch = _native.ptr(scratch)
    if adam is not None:
        # fused Adam/AMSGrad on the output rows: dict(param, exp_avg, exp_avg_sq, max_exp_avg_sq|None, hyper, beta1, beta2, eps, mirror|None)
        prm = adam["param"]
        for t in (prm, adam["exp_avg"], adam["exp_avg_sq"], adam.get("max_exp_avg_sq")):
            if t is not None and (t.dtype != torch.float32 or t.stride(1) != 1 or t.stride(0) != prm.stride(0)):
                raise RuntimeError("spmm: fused Adam tensors must be row-major fp32 with the same row pitch")
        a.adam_param, a.adam_exp_avg, a.adam_exp_avg_sq = prm.data_ptr(), adam["exp_avg"].data_ptr(), adam["exp_avg_sq"].data_ptr()
        a.adam_max_exp_avg_sq = _native.ptr(adam.get("max_exp_avg_sq"))
        a.adam_ld, a.adam_hyper_dev = prm.stride(0), adam["hyper"].data_ptr()
        a.adam_beta1, a.adam_beta2, a.adam_eps = adam.get("beta1", 0.9), adam.get("beta2", 0.999), adam.get("eps", 1e-8)
        a.adam_param_mirror_mc = adam.get("mirror")
        a.adam_mirror_rows = int(adam.get("mirror_rows") or 0)
    if plan.n_split_rows:
        a.slot_owner, a.split_counters = _native.ptr(plan.slot_owner), _native.ptr(plan.counters)
    a.B, a.ldb, a.b_dtype = B.data_ptr(), B.stride(0), _dt(B)
    if want_out:
        if out is None:
            out = torch.empty((n_out, F), dtype=out_dtype or torch.float32, device=B.device)
        if out.stride(1) != 1 or out.shape[0] < n_out or out.shape[1] < F:
            raise RuntimeError("spmm: bad `out` tensor")
        a.C, a.ldc, a.c_dtype = out.data_ptr(), out.stride(0), _dt(out)
    else:
        out = None
        a.C, a.ldc, a.c_dtype = None, 0, F32
    a.F = F
    a.c_row_offset = plan.row_begin
    a.bias = _native.ptr(bias)
    a.bias_len = int(bias.numel()) if bias is not None else 0
    a.act = act
    a.drop_mode, a.drop_p = drop_mode, float(drop_p)
    if keep_mask is not None:
        if keep_mask.dtype not in (torch.uint8, torch.bool) or keep_mask.stride(1) != 1:
            raise RuntimeError("spmm: keep_mask must be a row-major uint8/bool tensor")
        a.keep_mask, a.ldmask = keep_mask.data_ptr(), keep_mask.stride(0)
    a.philox_seed, a.philox_offset = philox_seed & (2**64 - 1), philox_offset & (2**64 - 1)
    a.philox_offset_dev = _native.ptr(philox_offset_dev)
    a.philox_row_offset = int(row_id_offset)
    if W_proj is not None:
        if W_proj.dtype != torch.float32 or not W_proj.is_contiguous() or W_proj.shape[0] != F:
            raise RuntimeError("spmm: W_proj must be a contiguous fp32 [F, n_proj] tensor")
        n_proj = int(W_proj.shape[1])
        if P is None:
            P = torch.zeros((n_out, pad4(n_proj)), dtype=torch.float32, device=B.device)
        a.W_proj, a.n_proj, a.P, a.ldp = W_proj.data_ptr(), n_proj, P.data_ptr(), P.stride(0)
    if scatter is not None:
        # dict(bases=int64 device tensor of peer-mapped base pointers, rows=rows per destination, row0=row offset there):
        # output row r goes to bases[r // rows] + (row0 + r % rows) * ldc (see tgcn_spmm_args.c_scatter_bases)
        a.c_scatter_bases, a.c_scatter_rows, a.c_scatter_row0 = scatter["bases"].data_ptr(), int(scatter["rows"]), int(scatter["row0"])
    if raw_slots is not None:
        # [n_slots, rows, >= F] fp32: partial rows added (slot order) to the first `rows` local rows before the epilogue
        if raw_slots.dtype != torch.float32 or raw_slots.dim() != 3 or raw_slots.stride(2) != 1 or raw_slots.shape[2] < F:
            raise RuntimeError("spmm: raw_slots must be an fp32 [n_slots, rows, >= F] tensor")
        a.raw_in, a.raw_ld, a.raw_stride = raw_slots.data_ptr(), raw_slots.stride(1), raw_slots.stride(0)
        a.n_raw, a.raw_rows = int(raw_slots.shape[0]), int(raw_slots.shape[1])
    if tc is not None:
        if graph is not tc.remainder:
            raise RuntimeError("spmm: with tc=..., graph must be the plan's remainder CSR")
        part = tc.buffers(F)[1]
        a.tc_part, a.tc_ld, a.tc_rank, a.tc_slot_ptr = part.data_ptr(), part.stride(0), tc.rank.data_ptr(), tc.slot_ptr.data_ptr()
    with torch.cuda.device(B.device):
        _native.check(lib.tgcn_spmm(C.byref(a), _stream()))
    return out, P


def spmm_hybrid(tc, B: torch.Tensor, *, F: Optional[int] = None, **kw):
    """The same operation as spmm() on the graph `tc` was built from, computed in two parts: the dense blocks on the
    tensor cores (tgcn_spmm_tc: operand pack + tcgen05 3xTF32 tiles -> partial rows), the remaining entries by the gather
    kernel, whose epilogue adds the partial rows before bias / activation / dropout / Adam.  Same keyword arguments as
    spmm() (bias, act, dropout, out, adam, ...); fp32 operands with 8 <= F <= 256 only."""
    _need_cuda(B)
    lib = _native.load()
    F = int(B.shape[1]) if F is None else int(F)
    if B.dtype != torch.float32 or B.stride(1) != 1:
        raise RuntimeError("spmm_hybrid: B must be a row-major fp32 matrix")
    bt, part = tc.buffers(F)
    cp = tc.c_struct()
    with torch.cuda.device(B.device):
        _native.check(lib.tgcn_spmm_tc(C.byref(cp), B.data_ptr(), B.stride(0), F, bt.data_ptr(), part.data_ptr(), part.stride(0),
                                       _stream()))
    return spmm(tc.remainder, B, F=F, tc=tc, **kw)


def masked_nll(Z: torch.Tensor, n_classes: int, y: torch.Tensor, mask: Optional[torch.Tensor], n_mask_total: int,
               *, want_grad: bool = True, dZ: Optional[torch.Tensor] = None, want_pred: bool = False,
               want_correct: bool = False, loss_out: Optional[torch.Tensor] = None,
               workspace: Optional[torch.Tensor] = None, pred: Optional[torch.Tensor] = None,
               correct: Optional[torch.Tensor] = None, want_partial: bool = False,
               partial: Optional[torch.Tensor] = None, dZ_mirror: Optional[int] = None,
               mask2: Optional[torch.Tensor] = None, correct2: Optional[torch.Tensor] = None):
    """Masked mean cross-entropy over rows of Z (+ gradient / argmax / #correct).  See tgcn_masked_nll.
    Returns dict(loss=[2] fp32 (mean nll, count), dZ, pred, correct, partial)."""
    _need_cuda(Z, y, mask, dZ, mask2, correct2)
    if correct2 is not None and (mask2 is None or mask2.dtype not in (torch.bool, torch.uint8) or not mask2.is_contiguous()):
        raise RuntimeError("masked_nll: correct2 needs a contiguous bool/uint8 mask2")
    lib = _native.load()
    n = int(Z.shape[0])
    if Z.dtype != torch.float32 or Z.stride(1) != 1:
        raise RuntimeError("masked_nll: Z must be a row-major fp32 tensor")
    if y.dtype != torch.int64 or not y.is_contiguous():
        raise RuntimeError("masked_nll: y must be a contiguous int64 tensor")
    if mask is not None and (mask.dtype not in (torch.bool, torch.uint8) or not mask.is_contiguous()):
        raise RuntimeError("masked_nll: mask must be a contiguous bool/uint8 tensor")
    dev = Z.device
    if loss_out is None:
        loss_out = torch.empty(2, dtype=torch.float32, device=dev)
    if partial is None and want_partial:
        partial = torch.empty(2, dtype=torch.float64, device=dev)
    if want_grad and dZ is None:
        dZ = torch.zeros((n, pad4(n_classes)), dtype=torch.float32, device=dev)
    if pred is None and want_pred:
        pred = torch.empty(n, dtype=torch.int32, device=dev)
    if correct is None and (want_correct or correct2 is not None):
        correct = torch.zeros(1, dtype=torch.int32, device=dev)
    need = 2 * ((n * 4 + 255) // 256 * 256) + 4096  # == tgcn_masked_nll_workspace_bytes(n)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _native.check(lib.tgcn_masked_nll(Z.data_ptr(), Z.stride(0), n, n_classes, y.data_ptr(), _native.ptr(mask),
                                          int(n_mask_total), loss_out.data_ptr(), _native.ptr(partial),
                                          _native.ptr(dZ) if want_grad else None, dZ.stride(0) if want_grad else 0,
                                          _native.ptr(pred), _native.ptr(correct), dZ_mirror,
                                          _native.ptr(mask2) if correct2 is not None else None, _native.ptr(correct2),
                                          workspace.data_ptr(), workspace.numel(), _stream()))
    return dict(loss=loss_out, dZ=dZ if want_grad else None, pred=pred, correct=correct, partial=partial, correct2=correct2)


def dense_bwd(G2: torch.Tensor, H1d: torch.Tensor, W2: torch.Tensor, dZ2: Optional[torch.Tensor], *,
              H: int, n_classes: int, act: int = ACT_NONE, drop_mode: int = DROP_NONE, drop_p: float = 0.0,
              keep_mask: Optional[torch.Tensor] = None, philox_seed: int = 0, philox_offset: int = 0,
              philox_offset_dev: Optional[torch.Tensor] = None, row_offset: int = 0, dZ1: Optional[torch.Tensor] = None, dz1_dtype: torch.dtype = torch.float32,
              want_dz1: bool = True, workspace: Optional[torch.Tensor] = None,
              dW2: Optional[torch.Tensor] = None, db_hidden: Optional[torch.Tensor] = None,
              db_out: Optional[torch.Tensor] = None, dZ1_mirror: Optional[int] = None, dZ1_mirror_rows: int = 0):
    """dW2 = H1d^T G2, db_out = colsum(dZ2), dZ1 = (G2 W2^T) * dropout' * act', db_hidden = colsum(dZ1)."""
    _need_cuda(G2, H1d, W2, dZ2, keep_mask, dZ1)
    lib = _native.load()
    dev = G2.device
    n = int(G2.shape[0])
    if W2.dtype != torch.float32 or not W2.is_contiguous() or tuple(W2.shape) != (H, n_classes):
        raise RuntimeError("dense_bwd: W2 must be a contiguous fp32 [H, C] tensor")
    a = _native.DenseBwdArgs()
    a.G2, a.ldg2 = G2.data_ptr(), G2.stride(0)
    a.H1d, a.ldh, a.h_dtype = H1d.data_ptr(), H1d.stride(0), _dt(H1d)
    a.W2 = W2.data_ptr()
    a.dZ2, a.lddz2 = (_native.ptr(dZ2), dZ2.stride(0)) if dZ2 is not None else (None, 0)
    a.n_rows, a.row_offset = n, row_offset
    a.H, a.C = H, n_classes
    a.act, a.drop_mode, a.drop_p = act, drop_mode, float(drop_p)
    if keep_mask is not None:
        a.keep_mask, a.ldmask = keep_mask.data_ptr(), keep_mask.stride(0)
    a.philox_seed, a.philox_offset = philox_seed & (2**64 - 1), philox_offset & (2**64 - 1)
    a.philox_offset_dev = _native.ptr(philox_offset_dev)
    if want_dz1:
        if dZ1 is None:
            width = pad4(H) if dz1_dtype == torch.float32 else pad8(H)
            dZ1 = torch.zeros((n, width), dtype=dz1_dtype, device=dev)
        a.dZ1, a.lddz1, a.dz1_dtype = dZ1.data_ptr(), dZ1.stride(0), _dt(dZ1)
        a.dZ1_mirror_mc = dZ1_mirror
        a.dZ1_mirror_rows = int(dZ1_mirror_rows)
    if dW2 is None:
        dW2 = torch.empty((H, n_classes), dtype=torch.float32, device=dev)
    if db_hidden is None and want_dz1:
        db_hidden = torch.empty(H, dtype=torch.float32, device=dev)
    if db_out is None and dZ2 is not None:
        db_out = torch.empty(n_classes, dtype=torch.float32, device=dev)
    a.dW2, a.db_hidden, a.db_out = dW2.data_ptr(), _native.ptr(db_hidden), _native.ptr(db_out)
    need = C.c_size_t(0)
    with torch.cuda.device(dev):
        _native.check(lib.tgcn_dense_bwd_workspace_bytes(H, n_classes, C.byref(need)))
        if workspace is None or workspace.numel() < need.value:
            workspace = torch.empty(need.value, dtype=torch.uint8, device=dev)
        _native.check(lib.tgcn_dense_bwd(C.byref(a), workspace.data_ptr(), workspace.numel(), _stream()))
    return dict(dW2=dW2, db_hidden=db_hidden, db_out=db_out, dZ1=dZ1 if want_dz1 else None, workspace=workspace)


def project(X: torch.Tensor, W: torch.Tensor, K: Optional[int] = None, out: Optional[torch.Tensor] = None,
            mirror: Optional[int] = None, bias: Optional[torch.Tensor] = None, drop_mode: int = DROP_NONE, drop_p: float = 0.0,
            keep_mask: Optional[torch.Tensor] = None, philox_seed: int = 0, philox_offset: int = 0,
            philox_offset_dev: Optional[torch.Tensor] = None, row_id_offset: int = 0,
            dropped_out: Optional[torch.Tensor] = None, mirror_rows: int = 0) -> torch.Tensor:
    """P = dropout(X[:, :K]) @ W (+ bias) (thin projection), P padded to a multiple of 4 columns.  With a dropout mode the
    keep decision of the SpMM epilogue is applied to X on load and dropout(X) is written to `dropped_out` (tgcn_project_ex)."""
    _need_cuda(X, W, out, bias, keep_mask, philox_offset_dev, dropped_out)
    lib = _native.load()
    K = int(W.shape[0]) if K is None else K
    M = int(W.shape[1])
    n = int(X.shape[0])
    if W.dtype != torch.float32 or not W.is_contiguous():
        raise RuntimeError("project: W must be contiguous fp32")
    if out is None:
        out = torch.zeros((n, pad4(M)), dtype=torch.float32, device=X.device)
    a = _native.ProjectArgs()
    a.X, a.ldx, a.x_dtype, a.n_rows, a.K = X.data_ptr(), X.stride(0), _dt(X), n, K
    a.W, a.M, a.bias = W.data_ptr(), M, _native.ptr(bias)
    a.P, a.ldp, a.P_mirror_mc, a.mirror_rows = out.data_ptr(), out.stride(0), mirror, int(mirror_rows)
    a.drop_mode, a.drop_p = drop_mode, float(drop_p)
    if keep_mask is not None:
        if keep_mask.dtype not in (torch.uint8, torch.bool) or keep_mask.stride(1) != 1:
            raise RuntimeError("project: keep_mask must be a row-major uint8/bool tensor")
        a.keep_mask, a.ldmask = keep_mask.data_ptr(), keep_mask.stride(0)
    a.philox_seed, a.philox_offset = philox_seed & (2**64 - 1), philox_offset & (2**64 - 1)
    a.philox_offset_dev, a.philox_row_offset = _native.ptr(philox_offset_dev), int(row_id_offset)
    if dropped_out is not None:
        if dropped_out.dtype != torch.float32 or dropped_out.stride(1) != 1 or dropped_out.shape[0] < n:
            raise RuntimeError("project: bad dropped_out tensor")
        a.Xd, a.ldxd = dropped_out.data_ptr(), dropped_out.stride(0)
    with torch.cuda.device(X.device):
        _native.check(lib.tgcn_project_ex(C.byref(a), _stream()))
    return out


def colsum(X: torch.Tensor, F: Optional[int] = None, out: Optional[torch.Tensor] = None,
           workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] = sum over rows of X[:, c] (deterministic).  See tgcn_colsum."""
    _need_cuda(X, out, workspace)
    lib = _native.load()
    F = int(X.shape[1]) if F is None else int(F)
    if X.dtype != torch.float32 or X.stride(1) != 1:
        raise RuntimeError("colsum: X must be a row-major fp32 matrix")
    if out is None:
        out = torch.empty(F, dtype=torch.float32, device=X.device)
    need = C.c_size_t(0)
    with torch.cuda.device(X.device):
        _native.check(lib.tgcn_colsum_workspace_bytes(F, C.byref(need)))
        if workspace is None or workspace.numel() < need.value:
            workspace = torch.empty(need.value, dtype=torch.uint8, device=X.device)
        _native.check(lib.tgcn_colsum(X.data_ptr(), X.stride(0), int(X.shape[0]), F, out.data_ptr(), workspace.data_ptr(),
                                      workspace.numel(), _stream()))
    return out


def dropout_apply(X: torch.Tensor, *, F: Optional[int] = None, out: Optional[torch.Tensor] = None, drop_mode: int = DROP_NONE,
                  drop_p: float = 0.0, keep_mask: Optional[torch.Tensor] = None, philox_seed: int = 0, philox_offset: int = 0,
                  philox_offset_dev: Optional[torch.Tensor] = None, row_id_offset: int = 0) -> torch.Tensor:
    """out = dropout(X[:, :F]) with the keep decision of the SpMM epilogue (see tgcn_dropout_apply)."""
    _need_cuda(X, out, keep_mask, philox_offset_dev)
    lib = _native.load()
    F = int(X.shape[1]) if F is None else int(F)
    if X.dtype != torch.float32 or X.dim() != 2 or X.stride(1) != 1:
        raise RuntimeError("dropout_apply: X must be a row-major fp32 matrix")
    if out is None:
        out = torch.empty((X.shape[0], F), dtype=torch.float32, device=X.device)
    if out.dtype != torch.float32 or out.stride(1) != 1 or out.shape[0] < X.shape[0] or out.shape[1] < F:
        raise RuntimeError("dropout_apply: bad `out` tensor")
    km, ldm = None, 0
    if keep_mask is not None:
        if keep_mask.dtype not in (torch.uint8, torch.bool) or keep_mask.stride(1) != 1:
            raise RuntimeError("dropout_apply: keep_mask must be a row-major uint8/bool tensor")
        km, ldm = keep_mask.data_ptr(), keep_mask.stride(0)
    with torch.cuda.device(X.device):
        _native.check(lib.tgcn_dropout_apply(X.data_ptr(), X.stride(0), out.data_ptr(), out.stride(0), int(X.shape[0]), F,
                                             drop_mode, float(drop_p), km, ldm, philox_seed & (2**64 - 1),
                                             philox_offset & (2**64 - 1), _native.ptr(philox_offset_dev), int(row_id_offset),
                                             _stream()))
    return out


def hier_forward(W1: torch.Tensor, n_nodes: int, n_vocab: int, Fdoc: torch.Tensor, out: Optional[torch.Tensor] = None):
    """XW = [I | F] @ W1 for X with hierarchy features on the document rows (text2graph.py:237-241)."""
    _need_cuda(W1, Fdoc, out)
    lib = _native.load()
    H = int(W1.shape[1])
    c_prev = int(Fdoc.shape[1])
    if W1.shape[0] != n_nodes + c_prev:
        raise RuntimeError(f"hier_forward: W1 has {W1.shape[0]} rows, expected n_nodes + c_prev = {n_nodes + c_prev}")
    if out is None:
        out = torch.zeros((n_nodes, pad4(H)), dtype=torch.float32, device=W1.device)
    with torch.cuda.device(W1.device):
        _native.check(lib.tgcn_hier_forward(W1.data_ptr(), W1.stride(0), n_nodes, n_vocab, Fdoc.data_ptr(), Fdoc.stride(0),
                                            c_prev, H, out.data_ptr(), out.stride(0), _stream()))
    return out


def hier_backward(G1: torch.Tensor, n_nodes: int, n_vocab: int, Fdoc: torch.Tensor, H: int, dW_tail: torch.Tensor):
    """dW1[N:, :] = F^T G1[doc rows]."""
    _need_cuda(G1, Fdoc, dW_tail)
    lib = _native.load()
    c_prev = int(Fdoc.shape[1])
    need = C.c_size_t(0)
    with torch.cuda.device(G1.device):
        _native.check(lib.tgcn_hier_backward_workspace_bytes(c_prev, H, C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=G1.device)
        if not dW_tail.is_contiguous():
            raise RuntimeError("hier_backward: dW_tail must be contiguous")
        _native.check(lib.tgcn_hier_backward(G1.data_ptr(), G1.stride(0), n_nodes, n_vocab, Fdoc.data_ptr(), Fdoc.stride(0),
                                             c_prev, H, dW_tail.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return dW_tail


def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
              max_exp_avg_sq: Optional[torch.Tensor], *, lr: float, beta1: float = 0.9, beta2: float = 0.999,
              eps: float = 1e-8, amsgrad: bool = False, step: int = 0, step_dev: Optional[torch.Tensor] = None,
              param_mirror: Optional[int] = None) -> None:
    """In-place fused Adam/AMSGrad update (torch.optim.Adam semantics, flat_amazon.py:89,106)."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, step_dev)
    lib = _native.load()
    for t in (param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise RuntimeError("adam_step: tensors must be contiguous fp32")
    with torch.cuda.device(param.device):
        _native.check(lib.tgcn_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                         _native.ptr(max_exp_avg_sq), param.numel(), lr, beta1, beta2, eps,
                                         1 if amsgrad else 0, int(step), _native.ptr(step_dev), param_mirror, _stream()))


def adam_step_small(params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq, *, lr: float, beta1: float = 0.9, beta2: float = 0.999,
                    eps: float = 1e-8, amsgrad: bool = False, step: int = 0, step_dev: Optional[torch.Tensor] = None) -> None:
    """One launch for up to 4 small parameter tensors (biases, W2): same arithmetic as adam_step."""
    lib = _native.load()
    k = len(params)
    _need_cuda(*params, *grads, *exp_avg, *exp_avg_sq, step_dev)
    for t in list(params) + list(grads) + list(exp_avg) + list(exp_avg_sq) + ([x for x in max_exp_avg_sq] if amsgrad else []):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("adam_step_small: tensors must be contiguous fp32")
    arr = lambda ts: (C.c_void_p * k)(*[t.data_ptr() for t in ts])
    sizes = (C.c_int64 * k)(*[p.numel() for p in params])
    with torch.cuda.device(params[0].device):
        _native.check(lib.tgcn_adam_step_small(k, arr(params), arr(grads), arr(exp_avg), arr(exp_avg_sq),
                                               arr(max_exp_avg_sq) if amsgrad else None, sizes, lr, beta1, beta2, eps,
                                               1 if amsgrad else 0, int(step), _native.ptr(step_dev), _stream()))


def adam_prepare(step_dev: torch.Tensor, hyper: torch.Tensor, lr: float, beta1: float = 0.9, beta2: float = 0.999) -> None:
    """step_dev += 1; hyper[0:2] = (lr / (1 - beta1^t), sqrt(1 - beta2^t)) for the SpMM's fused Adam epilogue."""
    lib = _native.load()
    with torch.cuda.device(step_dev.device):
        _native.check(lib.tgcn_adam_prepare(step_dev.data_ptr(), hyper.data_ptr(), lr, beta1, beta2, _stream()))


def increment_step(step_dev: torch.Tensor) -> None:
    lib = _native.load()
    with torch.cuda.device(step_dev.device):
        _native.check(lib.tgcn_increment_step(step_dev.data_ptr(), _stream()))


def count_mask(mask: torch.Tensor) -> torch.Tensor:
    _need_cuda(mask)
    lib = _native.load()
    out = torch.empty(1, dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        _native.check(lib.tgcn_count_mask(mask.data_ptr(), mask.numel(), out.data_ptr(), _stream()))
    return out


def cast_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(src, out)
    lib = _native.load()
    if not src.is_contiguous() or src.dtype != torch.float32:
        raise RuntimeError("cast_bf16: src must be contiguous fp32")
    if out is None:
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        _native.check(lib.tgcn_cast_f32_to_bf16(src.data_ptr(), out.data_ptr(), src.numel(), _stream()))
    return out
