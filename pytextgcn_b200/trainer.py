"""Fused full-batch trainer: the reference's epoch loop (flat_amazon.py:99-117) with every
device-side step routed through the C ABI and captured in CUDA graphs.

One reference epoch =
    train step : gcn(g)[train_mask] -> CrossEntropyLoss(mean) -> backward -> Adam/AMSGrad step
                 (flat_amazon.py:100-106)
    eval       : gcn.eval() forward, val loss, argmax / accuracy  (flat_amazon.py:107-116)
Here the train step is 8 kernels families (2 forward SpMMs with fused bias/dropout/projection,
masked NLL + gradient, 2 backward SpMMs, the dense dW/db/dZ1 pass, Adam) on static buffers;
parameters are updated in place, so the wrapped `GCN` stays an ordinary nn.Module
(state_dict / th.save keep working, flat_amazon.py:126-128).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .graph import GraphCSR, get_graph
from .models import GCN, decode_features


class TextGCNTrainer:
    def __init__(self, gcn: GCN, g, lr: float = 0.05, amsgrad: bool = True, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_cuda_graph: bool = True, seed: int = 0, chunk_nnz: Optional[int] = None,
                 graph: Optional[GraphCSR] = None, assume_symmetric: bool = False, eval_mode: str = "layered",
                 fuse_adam: bool = True, keep_w1_grad: bool = True, share_h1: bool = True, restrict_rows: bool = True,
                 tensor_cores: Optional[bool] = None, tc_min_density: float = 0.05):
        if len(gcn.layers) != 2:
            raise NotImplementedError("TextGCNTrainer fuses the 2-layer TextGCN (the configuration of every reference script)")
        l0, l1 = gcn.layers
        if not l0.weight.is_cuda:
            raise RuntimeError("TextGCNTrainer needs the module and the graph on a CUDA device (no CPU path)")
        self.gcn, self.g = gcn, g
        self.dev = l0.weight.device
        self.n = int(g.x.shape[0])
        self.feat = decode_features(g.x, getattr(g, "n_vocab", None))
        if self.feat is None:
            raise RuntimeError("TextGCNTrainer expects featureless input x = I or [I | F] (text2graph.py:226-246)")
        self.graph = graph if graph is not None else get_graph(g.edge_index, g.edge_attr, self.n, holder=g)
        if assume_symmetric:
            self.graph._symmetric = True
        self.graph_t = self.graph.transpose()
        kw = {} if chunk_nnz is None else dict(chunk_nnz=chunk_nnz)
        self.plan = self.graph.plan(**kw)
        self.plan_t = self.plan if self.graph_t is self.graph else self.graph_t.plan(**kw)
        # The loss and the metrics only ever read the logits of MASKED rows (gcn(g)[g.train_mask], logits[g.val_mask],
        # flat_amazon.py:101,110-112), and the loss gradient dZ2 is zero outside the training rows.  With restrict_rows
        # the class-wide propagations skip what nothing reads: the forward computes the logits of the rows selected by
        # any of the masks (train | val | test -- the document rows), the backward multiplies only the columns that
        # carry a non-zero gradient.  Same sums without the terms that are exactly zero; `logits`/`pred` are then
        # defined on masked rows only (eval_step(full=True) computes every row).
        self.restrict_rows = bool(restrict_rows)
        self.lr, self.amsgrad, self.betas, self.eps = float(lr), bool(amsgrad), betas, float(eps)
        # Hybrid hidden-wide propagation (csrc/spmm_tc.cu): the dense blocks of A_hat on the tensor cores (3xTF32, fp32
        # accuracy), the rest gathered as before.  None = use it when the graph has dense blocks worth it.
        self.tc = self.tc_t = None
        Hh = int(gcn.layers[0].weight.shape[1])
        if tensor_cores and not (Hh % 4 == 0 and 64 <= Hh <= 256):
            raise RuntimeError("tensor_cores=True needs a hidden width that is a multiple of 4 in [64, 256]")
        # auto mode: graphs big enough to matter whose gathered operand (N x H fp32) is L2-resident -- beyond that the gather
        # kernel is bound by L2 misses on the non-hub columns, which the tensor-core part does not relieve (measured on the
        # 1.2 M-node graph: 11.7 ms hybrid vs 11.4 ms gather, profiles/r02_hybrid_spmm.json)
        auto = tensor_cores is None and self.graph.nnz >= 200_000 and self.n * ((Hh + 15) // 16 * 16) * 8 <= (104 << 20)
        if (tensor_cores or auto) and Hh % 4 == 0 and 64 <= Hh <= 256:
            from .tc_plan import build_tc_plan
            n_sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
            tc = build_tc_plan(self.graph, min_density=tc_min_density, n_sms=n_sms, width=Hh)
            if tc is not None and (tensor_cores or tc.nnz_dense >= 0.15 * self.graph.nnz):
                self.tc = tc
                self.tc_t = tc if self.graph_t is self.graph else build_tc_plan(self.graph_t, min_density=tc_min_density,
                                                                                n_sms=n_sms, width=Hh)

        self.act = ops.ACT_RELU if gcn.apply_activation else ops.ACT_NONE
        self.p = float(gcn.dropout)
        self.seed = int(seed)
        self.use_cuda_graph = use_cuda_graph
        self.H, self.C = int(l0.weight.shape[1]), int(l1.weight.shape[1])
        self.in_ch = int(l0.weight.shape[0])
        if self.H % 4 != 0:
            raise NotImplementedError("TextGCNTrainer needs a hidden width that is a multiple of 4 "
                                      "(reference widths: 32, 64, 100, 200, 256)")
        self.Cp = ops.pad4(self.C)
        n, H, Cp, dev = self.n, self.H, self.Cp, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        self.params = [l0.weight, l0.bias, l1.weight, l1.bias]
        for p_ in self.params:
            if not p_.is_contiguous():
                raise RuntimeError("parameters must be contiguous")
        self.grads = [torch.zeros_like(p_.data) for p_ in self.params]
        self.exp_avg = [torch.zeros_like(p_.data) for p_ in self.params]
        self.exp_avg_sq = [torch.zeros_like(p_.data) for p_ in self.params]
        self.max_exp_avg_sq = [torch.zeros_like(p_.data) for p_ in self.params] if amsgrad else [None] * 4
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.adam_hyper = torch.zeros(2, dtype=torch.float32, device=dev)   # lr/(1-b1^t), sqrt(1-b2^t) of the current step
        # W1's Adam update can run in the epilogue of the SpMM that produces dW1 (rows of dW1 never leave registers
        # unless keep_w1_grad asks for the gradient buffer too); not with hierarchy features (W1 has extra rows)
        self.fuse_adam = bool(fuse_adam) and self.feat.Fdoc is None
        self.keep_w1_grad = bool(keep_w1_grad)
        self.XW = torch.zeros((n, H), **f32) if self.feat.Fdoc is not None else None
        self.H1d = torch.empty((n, H), **f32)
        # A_hat (X W1) + b1 of the eval forward of epoch k IS the pre-dropout hidden activation of the training forward
        # of epoch k+1: W1/b1 do not change between them (flat_amazon.py:100-110: step -> eval -> next step) and
        # F.dropout follows the product (models.py:20-23).  With share_h1 the eval pass keeps it in `H1` and the next
        # train step only applies its dropout mask (tgcn_dropout_apply) instead of repeating the hidden-wide SpMM:
        # bit-identical activations, one wide propagation less per epoch.  `_h1_key` names the parameter state H1 holds.
        self.share_h1 = bool(share_h1)
        self.H1 = torch.empty((n, H), **f32) if (self.share_h1 and self.p > 0.0) else self.H1d
        self._h1_key = None
        # Propagate-first order for layer 2 when there are more classes than hidden units (perlevel_dbpedia.py: 219 / 70
        # classes, hidden 32): Z2 = (A_hat H1d) W2 + b2 instead of A_hat (H1d W2) + b2, and in the backward
        # dH1d = A_hat^T (dZ2 W2^T), dW2 = (A_hat H1d)^T dZ2 -- the same linear maps re-associated, so both class-wide
        # propagations move hidden-wide rows (7x fewer gathered bytes at 219 vs 32).  Identity activation only.
        self.propagate_first = bool(Cp > H and self.act == ops.ACT_NONE)
        if self.propagate_first:
            self.U = torch.zeros((n, H), **f32)            # A_hat H1d
            self.T = torch.zeros((n, H), **f32)            # dZ2 W2^T
            self.W2t = torch.zeros((self.C, H), **f32)
            self._cs_ws = torch.empty(4096 * H * 4, dtype=torch.uint8, device=dev)     # colsum partials (>= 4 CTAs per SM)
        self.P = torch.zeros((n, Cp), **f32)
        self.Z2 = torch.zeros((n, Cp), **f32)
        self.dZ2 = torch.zeros((n, Cp), **f32)
        self.G2 = torch.zeros((n, Cp), **f32)
        self.dZ1 = torch.zeros((n, H), **f32)
        self.loss_train = torch.zeros(2, **f32)
        self.loss_val = torch.zeros(2, **f32)
        self.loss_tr_eval = torch.zeros(2, **f32)
        self.pred = torch.zeros(n, dtype=torch.int32, device=dev)
        self.correct_val = torch.zeros(1, dtype=torch.int32, device=dev)
        self.correct_train = torch.zeros(1, dtype=torch.int32, device=dev)
        self.hier_tail = torch.empty((self.in_ch - n, H), **f32) if self.feat.Fdoc is not None else None
        self._nll_ws = torch.empty(2 * ((n * 4 + 255) // 256 * 256) + 4096, dtype=torch.uint8, device=dev)
        self._db_ws = None
        self.logits = self.Z2[:, :self.C]
        self.Q = torch.zeros((n, Cp), **f32)          # collapsed eval: X W1 W2
        self.Tc = torch.zeros((n, Cp), **f32)         # collapsed eval: A_hat Q + 1 (b1^T W2)
        self.c_row = torch.zeros((1, Cp), **f32)
        self.eval_mode = "layered"
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._warm: Dict[str, int] = {}
        self.set_masks(g.y, g.train_mask, getattr(g, "val_mask", None), getattr(g, "test_mask", None))
        self.launches_per_train_step = 0
        self.launches_per_train_step_reuse = 0
        self.launches_per_eval = 0
        self.set_eval_mode(eval_mode)

    def set_eval_mode(self, mode: str) -> None:
        """"layered": the reference's operation order, A (A (X W1) + b1) W2 ... evaluated layer by layer.
        "collapsed": the same function re-associated as A (A (X W1 W2) + 1 (b1^T W2)) + b2 -- valid because
        the reference GCN applies no activation (models.py:22) and dropout is off in eval; it replaces
        the hidden-wide propagation of the eval forward by a classes-wide one (10x fewer gathered bytes
        at hidden 200 / 20 classes).  Logits agree to fp32 rounding (tests/test_gpu_train.py)."""
        if mode not in ("layered", "collapsed"):
            raise ValueError("eval_mode must be 'layered' or 'collapsed'")
        if mode == "collapsed" and self.act != ops.ACT_NONE:
            raise ValueError("collapsed eval needs the activation-free reference model (apply_activation=False)")
        if mode != self.eval_mode:
            self._graphs.pop("eval", None)
            self._warm.pop("eval", None)
        self.eval_mode = mode

    # ---- labels / masks (per-call inputs of the loss, flat_amazon.py:101-102,110) ----
    def set_masks(self, y: torch.Tensor, train_mask: torch.Tensor, val_mask: Optional[torch.Tensor],
                  test_mask: Optional[torch.Tensor] = None) -> None:
        """Copies labels and masks into static device buffers (so captured graphs stay valid), recounts, and
        rebuilds the mask-dependent work lists when a mask changed.  One host sync."""
        dev = self.dev
        first = not hasattr(self, "y")
        if first:
            self.y = torch.empty(self.n, dtype=torch.int64, device=dev)
            self.train_mask = torch.zeros(self.n, dtype=torch.bool, device=dev)
            self.val_mask = torch.zeros(self.n, dtype=torch.bool, device=dev)
            self.test_mask = torch.zeros(self.n, dtype=torch.bool, device=dev)
        old = None if first else (self.train_mask.clone(), self.val_mask.clone(), self.test_mask.clone())
        self.y.copy_(y, non_blocking=True)
        self.train_mask.copy_(train_mask, non_blocking=True)
        if val_mask is not None:
            self.val_mask.copy_(val_mask, non_blocking=True)
        if test_mask is not None:
            self.test_mask.copy_(test_mask, non_blocking=True)
        n_train = int(self.train_mask.sum().item())
        n_val = int(self.val_mask.sum().item())
        # labels of the rows the loss reads must be class ids (CrossEntropyLoss raises on anything else,
        # flat_amazon.py:102); rows outside the masks may hold -1 (perlabel_amazon.py:108-109) and are never read
        used = self.train_mask | self.val_mask
        bad = used & ((self.y < 0) | (self.y >= self.C))
        if bool(bad.any().item()):
            raise RuntimeError(f"labels of masked rows must lie in [0, {self.C}): found {int(bad.sum().item())} outside")
        changed = first or not all(torch.equal(a, b) for a, b in zip(old, (self.train_mask, self.val_mask, self.test_mask)))
        if getattr(self, "n_train", n_train) != n_train or getattr(self, "n_val", n_val) != n_val or (changed and self.restrict_rows):
            self._graphs = {}      # the divisor / the work lists are baked into the captured launches
            self._warm = {}
        self.n_train, self.n_val = n_train, n_val
        if n_train == 0:
            raise RuntimeError("train_mask selects no rows")
        if changed:
            self._build_restricted()

    def _build_restricted(self) -> None:
        if not self.restrict_rows:
            self.plan_z2, self.graph_g2, self.plan_g2 = self.plan, self.graph_t, self.plan_t
            return
        rows = self.train_mask | self.val_mask | self.test_mask
        self.plan_z2 = self.graph.plan_for_rows(rows, self.plan)                 # logits of the masked rows
        self.graph_g2 = self.graph_t.select_columns(self.train_mask)             # A_hat^T restricted to the columns where dZ2 != 0
        self.plan_g2 = self.graph_g2.plan_nonempty()
        self.G2.zero_()                                                          # rows without such a column stay zero
        self.Z2.zero_()
        self.dZ1.zero_()
        if self.propagate_first:
            self.U.zero_()

    def update_inputs(self, y_host: torch.Tensor, train_mask_host: torch.Tensor, val_mask_host: torch.Tensor) -> None:
        """Per-epoch inputs of the loss from (pinned) HOST memory: asynchronous H2D copies into the static buffers.
        The masks are compared with the previous call's on the host (a few microseconds); only a changed mask goes
        through set_masks (recount + rebuild of the restricted work lists)."""
        last = getattr(self, "_host_masks", None)
        if last is not None and torch.equal(last[0], train_mask_host) and torch.equal(last[1], val_mask_host):
            self.y.copy_(y_host, non_blocking=True)
            self.train_mask.copy_(train_mask_host, non_blocking=True)
            self.val_mask.copy_(val_mask_host, non_blocking=True)
            return
        self.set_masks(y_host, train_mask_host, val_mask_host)
        self._host_masks = (train_mask_host.clone(), val_mask_host.clone())

    # ---- the step bodies (eager; captured once warmed up) ----
    def _forward_collapsed(self, full: bool = False) -> None:
        l0, l1 = self.gcn.layers
        W1, b1, W2, b2 = l0.weight.data, l0.bias.data, l1.weight.data, l1.bias.data
        if self.feat.Fdoc is not None:
            ops.hier_forward(W1, self.n, self.feat.n_vocab, self.feat.Fdoc, out=self.XW)
            B1 = self.XW
        else:
            B1 = W1[:self.n]
        ops.project(B1, W2, K=self.H, out=self.Q)                       # Q = (X W1) W2
        ops.project(b1.view(1, self.H), W2, K=self.H, out=self.c_row)    # c = b1^T W2
        ops.spmm(self.graph, self.Q, F=self.Cp, plan=self.plan, out=self.Tc, bias=self.c_row[0, :self.C])
        ops.spmm(self.graph, self.Tc, F=self.Cp, plan=self.plan if full else self.plan_z2, out=self.Z2, bias=b2)

    def _wide_spmm(self, transposed: bool, B: torch.Tensor, **kw):
        """Hidden-wide propagation A_hat B (or A_hat^T B) with the fused epilogue `kw`: hybrid when a plan exists."""
        tc = self.tc_t if transposed else self.tc
        if tc is not None and kw.get("W_proj") is None:
            return ops.spmm_hybrid(tc, B, F=self.H, plan=tc.remainder.plan(), **kw)
        graph, plan = (self.graph_t, self.plan_t) if transposed else (self.graph, self.plan)
        return ops.spmm(graph, B, F=self.H, plan=plan, **kw)

    def _param_key(self):
        l0 = self.gcn.layers[0]
        return (l0.weight._version, l0.bias._version, l0.weight.data_ptr(), l0.bias.data_ptr())

    def invalidate_cache(self) -> None:
        """Forget the shared hidden activation (call after changing W1/b1 through raw pointers or `.data`;
        in-place torch ops on the Parameters such as load_state_dict are detected through their version counters)."""
        self._h1_key = None

    def _forward(self, training: bool, reuse_h1: bool = False, full: bool = False) -> None:
        l0, l1 = self.gcn.layers
        W1, b1, W2, b2 = l0.weight.data, l0.bias.data, l1.weight.data, l1.bias.data
        if not training and self.eval_mode == "collapsed":
            self._forward_collapsed(full)
            return
        drop = training and self.p > 0.0
        dkw = dict(drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed,
                   philox_offset=0, philox_offset_dev=self.step_dev if drop else None)
        fuse = self.C <= ops.FUSED_PROJ_MAX_CLASSES and not reuse_h1
        if reuse_h1:
            # H1 = act(A_hat (X W1) + b1) of the preceding eval pass, same W1/b1: only the dropout mask is applied --
            # inside the projection kernel, which also writes the dropped activation for the backward pass
            if drop and not self.propagate_first:
                ops.project(self.H1, W2, K=self.H, out=self.P, dropped_out=self.H1d, **dkw)
                ops.spmm(self.graph, self.P, F=self.Cp, plan=self.plan if full else self.plan_z2, out=self.Z2, bias=b2)
                return
            if drop:
                ops.dropout_apply(self.H1, F=self.H, out=self.H1d, **dkw)
            h_out = self.H1d
        else:
            if self.feat.Fdoc is not None:
                ops.hier_forward(W1, self.n, self.feat.n_vocab, self.feat.Fdoc, out=self.XW)
                B1 = self.XW
            else:
                B1 = W1[:self.n]
            h_out = self.H1d if training else self.H1
            self._wide_spmm(False, B1, out=h_out, bias=b1, act=self.act,
                            W_proj=W2 if fuse else None, P=self.P if fuse else None, **dkw)
        if self.propagate_first:
            ops.spmm(self.graph, h_out, F=self.H, plan=self.plan if full else self.plan_z2, out=self.U)
            ops.project(self.U, W2, K=self.H, out=self.Z2, bias=b2)
            return
        if not fuse:
            ops.project(h_out, W2, K=self.H, out=self.P)
        ops.spmm(self.graph, self.P, F=self.Cp, plan=self.plan if full else self.plan_z2, out=self.Z2, bias=b2)

    def _train_body(self, reuse_h1: bool = False) -> None:
        l0, l1 = self.gcn.layers
        W2 = l1.weight.data
        self._forward(True, reuse_h1)
        ops.masked_nll(self.Z2, self.C, self.y, self.train_mask, self.n_train, want_grad=True, dZ=self.dZ2,
                       loss_out=self.loss_train, workspace=self._nll_ws)
        drop = self.p > 0.0
        dkw = dict(drop_mode=ops.DROP_PHILOX if drop else ops.DROP_NONE, drop_p=self.p, philox_seed=self.seed, philox_offset=0,
                   philox_offset_dev=self.step_dev if drop else None)
        if self.propagate_first:
            self.W2t.copy_(W2.t())
            ops.project(self.dZ2, self.W2t, K=self.C, out=self.T)                            # T = dZ2 W2^T (zero off the train rows)
            ops.spmm(self.graph_g2, self.T, F=self.H, plan=self.plan_g2, out=self.dZ1, **dkw)  # dZ1 = dropout'(A_hat^T T)
            r = ops.dense_bwd(self.dZ2, self.U, W2, self.dZ2, H=self.H, n_classes=self.C, want_dz1=False,
                              workspace=self._db_ws, dW2=self.grads[2], db_out=self.grads[3])   # dW2 = U^T dZ2, db2 = colsum(dZ2)
            ops.colsum(self.dZ1, F=self.H, out=self.grads[1], workspace=self._cs_ws)
        else:
            ops.spmm(self.graph_g2, self.dZ2, F=self.Cp, plan=self.plan_g2, out=self.G2)
            r = ops.dense_bwd(self.G2, self.H1d, W2, self.dZ2, H=self.H, n_classes=self.C, act=self.act,
                              dZ1=self.dZ1, workspace=self._db_ws, dW2=self.grads[2], db_hidden=self.grads[1],
                              db_out=self.grads[3], **dkw)
        self._db_ws = r["workspace"]
        kw = dict(lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, amsgrad=self.amsgrad, step_dev=self.step_dev)
        if self.fuse_adam:
            ops.adam_prepare(self.step_dev, self.adam_hyper, self.lr, self.betas[0], self.betas[1])     # step += 1
            self._wide_spmm(True, self.dZ1, out=self.grads[0] if self.keep_w1_grad else None,
                            want_out=self.keep_w1_grad,
                            adam=dict(param=self.params[0].data, exp_avg=self.exp_avg[0], exp_avg_sq=self.exp_avg_sq[0],
                                      max_exp_avg_sq=self.max_exp_avg_sq[0], hyper=self.adam_hyper, beta1=self.betas[0],
                                      beta2=self.betas[1], eps=self.eps))
        else:
            self._wide_spmm(True, self.dZ1, out=self.grads[0])
            if self.feat.Fdoc is not None:
                ops.hier_backward(self.grads[0], self.n, self.feat.n_vocab, self.feat.Fdoc, self.H, self.hier_tail)
                self.grads[0][self.n:].copy_(self.hier_tail)
            ops.increment_step(self.step_dev)
            ops.adam_step(self.params[0].data, self.grads[0], self.exp_avg[0], self.exp_avg_sq[0], self.max_exp_avg_sq[0], **kw)
        ops.adam_step_small([p_.data for p_ in self.params[1:]], self.grads[1:], self.exp_avg[1:], self.exp_avg_sq[1:],
                            self.max_exp_avg_sq[1:], **kw)

    def _eval_body(self, full: bool = False) -> None:
        """eval forward, val loss, argmax, #correct on the val and train rows
        (flat_amazon.py:107-114 without the D2H copies)."""
        self._forward(False, full=full)
        if self.n_val > 0:      # one pass: val loss, argmax, #correct on the val rows and (mask2) on the train rows
            ops.masked_nll(self.Z2, self.C, self.y, self.val_mask, self.n_val, want_grad=False,
                           loss_out=self.loss_val, workspace=self._nll_ws, pred=self.pred, correct=self.correct_val,
                           mask2=self.train_mask, correct2=self.correct_train)
        else:
            ops.masked_nll(self.Z2, self.C, self.y, self.train_mask, self.n_train, want_grad=False,
                           loss_out=self.loss_tr_eval, workspace=self._nll_ws, pred=self.pred, correct=self.correct_train)

    # ---- graph capture plumbing ----
    def _run(self, name: str, body) -> None:
        from . import _native
        lib = _native.load()
        if not self.use_cuda_graph:
            c0 = lib.tgcn_launch_count()
            body()
            self._set_launches(name, int(lib.tgcn_launch_count() - c0))
            return
        gph = self._graphs.get(name)
        if gph is not None:
            gph.replay()
            return
        warm = self._warm.get(name, 0)
        if warm < 2:                      # eager warm-up (sets func attributes, sizes workspaces)
            c0 = lib.tgcn_launch_count()
            body()
            self._set_launches(name, int(lib.tgcn_launch_count() - c0))   # exact kernel count of this step
            self._warm[name] = warm + 1
            return
        torch.cuda.synchronize(self.dev)
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            body()
        self._graphs[name] = gph
        gph.replay()

    def _set_launches(self, name: str, k: int) -> None:
        if name == "train":
            self.launches_per_train_step = k
        elif name == "train_reuse":
            self.launches_per_train_step_reuse = k
        else:
            self.launches_per_eval = k

    # ---- public API ----
    def train_step(self) -> torch.Tensor:
        """One full-batch training step.  Returns the device tensor [mean train loss, #rows]."""
        self.gcn.train()
        if self.share_h1 and self._h1_key is not None and self._h1_key == self._param_key():
            self._run("train_reuse", lambda: self._train_body(True))
        else:
            self._run("train", self._train_body)
        self._h1_key = None          # W1/b1 have just been updated ...
        if hasattr(self.gcn, "invalidate_cache"):
            self.gcn.invalidate_cache()   # ... through raw pointers: the module's own cache cannot see that
        return self.loss_train

    def eval_step(self, full: bool = False) -> Dict[str, torch.Tensor]:
        """Eval forward + val loss + on-device argmax/accuracy counts (device tensors, no sync).  `logits` / `pred`
        hold the rows selected by any of the masks; full=True computes the logits of every row (word rows included)."""
        self.gcn.eval()
        if full and self.restrict_rows:
            self._run("eval_full", lambda: self._eval_body(True))
        else:
            self._run("eval", self._eval_body)
        if self.share_h1 and self.eval_mode == "layered":
            self._h1_key = self._param_key()      # H1 now holds act(A_hat (X W1) + b1) for the current W1/b1
        return dict(logits=self.logits, val_loss=self.loss_val, pred=self.pred, correct_val=self.correct_val,
                    correct_train=self.correct_train)

    def epoch(self) -> Dict[str, float]:
        """One reference epoch; the only host sync is the final read-back (like loss.item(),
        flat_amazon.py:115)."""
        self.train_step()
        self.eval_step()
        vals = torch.cat([self.loss_train[:1], self.loss_val[:1],
                          self.correct_train.float() / max(self.n_train, 1),
                          self.correct_val.float() / max(self.n_val, 1)]).cpu().tolist()
        return dict(loss=vals[0], val_loss=vals[1], acc_train=vals[2], acc_val=vals[3])

    @property
    def step(self) -> int:
        return int(self.step_dev.item())
