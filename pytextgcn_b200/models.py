"""Drop-in `GCN` / `GCNConv` modules over the sm_100a kernels.

Mirrors the reference's public surface for the hot path:
  * `GCN(in_channels, out_channels, n_gcn=2, n_hidden_gcn=64, activation=nn.ReLU, dropout=0.5)`
    and `forward(g) -> logits[N, out_channels]` -- textgcn/lib/models.py:6-25;
  * `GCNConv(in, out, add_self_loops=True)` called as `layer(x, edge_index, edge_weight)` --
    models.py:11-15,20 ([PyG-1.6.3] torch_geometric.nn.GCNConv), parameters `weight (in,out)`,
    `bias (out)`, Glorot / zeros init.
As in the reference, `self.activation` is constructed but NOT applied (models.py:22 is
commented out); pass `apply_activation=True` for the TextGCN-paper ReLU variant.

Everything below `forward` runs through the C ABI (pytextgcn_b200/csrc); there is no eager
or CPU fallback -- CPU tensors raise RuntimeError.
"""
from __future__ import annotations

import math
import weakref
from typing import Optional

import torch
from torch import nn

from . import ops
from .graph import GraphCSR, get_graph


# --------------------------------------------------------------------------------------
# feature-matrix decoding:  x = I_N  or  [I_N | F]  (text2graph.py:226-246)
# --------------------------------------------------------------------------------------
class _FeatInfo:
    __slots__ = ("n_nodes", "n_cols", "Fdoc", "n_vocab", "identity")

    def __init__(self, n_nodes, n_cols, Fdoc, n_vocab, identity=True):
        self.n_nodes, self.n_cols, self.Fdoc, self.n_vocab, self.identity = n_nodes, n_cols, Fdoc, n_vocab, identity


_FEAT_CACHE: list = []


def decode_features(x: torch.Tensor, n_vocab: Optional[int] = None) -> Optional[_FeatInfo]:
    """Recognise the featureless input X = I_N (sparse COO, text2graph.py:234,243-246) or
    X = [I_N | F] with F on the document rows (text2graph.py:237-241) and return F as a dense
    [n_docs, C_prev] matrix.  Returns None when x is a general matrix.  One-off per tensor."""
    for e in list(_FEAT_CACHE):
        t = e[0]()
        if t is None:
            _FEAT_CACHE.remove(e)
        elif t is x and e[1] == x._version:
            return e[2]
    info = None
    if x.is_sparse and x.dim() == 2 and x.shape[1] >= x.shape[0]:
        n, m = int(x.shape[0]), int(x.shape[1])
        xc = x if x.is_coalesced() else x.coalesce()
        idx, vals = xc.indices(), xc.values()
        r, c = idx[0], idx[1]
        diag = (r == c)
        ok = int(diag.sum().item()) == n and bool((vals[diag] == 1).all().item())
        rest = ~diag
        if ok and m == n and not bool(rest.any().item()):
            info = _FeatInfo(n, m, None, n_vocab if n_vocab is not None else 0)
        elif ok and m > n:
            rr, cc, vv = r[rest], c[rest], vals[rest]
            if rr.numel() == 0:
                nv = n if n_vocab is None else n_vocab
            else:
                nv = int(rr.min().item()) if n_vocab is None else int(n_vocab)
            if bool((cc >= n).all().item()) and (rr.numel() == 0 or int(rr.min().item()) >= nv):
                Fdoc = torch.zeros((n - nv, m - n), dtype=torch.float32, device=x.device)
                Fdoc.index_put_((rr - nv, cc - n), vv.to(torch.float32))
                info = _FeatInfo(n, m, Fdoc, nv)
    _FEAT_CACHE.append((weakref.ref(x), x._version, info))
    if len(_FEAT_CACHE) > 16:
        _FEAT_CACHE.pop(0)
    return info


def _glorot_(w: torch.Tensor) -> None:
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    with torch.no_grad():
        w.uniform_(-a, a)


class DropoutSpec:
    """How the fused epilogue drops: nothing, a caller-supplied keep-mask (parity tests), or
    Philox regenerated in the backward pass (default in training; F.dropout in the reference,
    models.py:23)."""
    __slots__ = ("mode", "p", "mask", "seed", "offset")

    def __init__(self, mode=ops.DROP_NONE, p=0.0, mask=None, seed=0, offset=0):
        self.mode, self.p, self.mask, self.seed, self.offset = mode, float(p), mask, int(seed), int(offset)


_NO_DROP = DropoutSpec()


def _identity_operand(W: torch.Tensor, feat: _FeatInfo) -> torch.Tensor:
    """X @ W for X = I_N or [I_N | F] as a [N, pad4(H)] fp32 operand (zero-copy when possible)."""
    n, H = feat.n_nodes, int(W.shape[1])
    if feat.Fdoc is None:
        if H % 4 == 0 and W.stride(0) % 4 == 0 and W.stride(1) == 1 and W.data_ptr() % 16 == 0:
            return W[:n]
        out = torch.zeros((n, ops.pad4(H)), dtype=torch.float32, device=W.device)
        out[:, :H].copy_(W[:n])
        return out
    return ops.hier_forward(W.contiguous(), n, feat.n_vocab, feat.Fdoc)


# --------------------------------------------------------------------------------------
# fused 2-layer TextGCN:  logits = A (drop(act(A (X W1) + b1)) W2) + b2
# --------------------------------------------------------------------------------------
class _HiddenCache:
    """Pre-dropout hidden activation act(A_hat (X W1) + b1) of the last eval forward, keyed by the state of
    W1/b1 (version counters + storage) and the graph.  The reference loop runs step -> eval -> next step
    (flat_amazon.py:100-110): the eval forward of epoch k and the training forward of epoch k+1 see the same
    W1/b1, and F.dropout comes after the product (models.py:20-23), so the next training forward only has to
    apply its dropout mask -- bit-identical to recomputing the propagation."""
    __slots__ = ("key", "h1")

    def __init__(self):
        self.key, self.h1 = None, None


class _GCN2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W1, b1, W2, b2, graph: GraphCSR, feat: _FeatInfo, act: int, drop: DropoutSpec,
                cache: Optional[_HiddenCache] = None, cache_key=None):
        n = graph.n_nodes
        H, C = int(W1.shape[1]), int(W2.shape[1])
        Hp, Cp = ops.pad4(H), ops.pad4(C)
        if drop.mode == ops.DROP_MASK and int(drop.mask.shape[1]) != Hp:
            m = torch.zeros((n, Hp), dtype=torch.uint8, device=W1.device)     # pad the caller's keep-mask
            m[:, :drop.mask.shape[1]].copy_(drop.mask)
            drop = DropoutSpec(drop.mode, drop.p, m, drop.seed, drop.offset)
        W2c = W2.contiguous()
        if Hp != H:
            W2p = torch.zeros((Hp, C), dtype=torch.float32, device=W2.device)
            W2p[:H].copy_(W2c)
        else:
            W2p = W2c
        P = torch.zeros((n, Cp), dtype=torch.float32, device=W1.device) if Cp != C else \
            torch.empty((n, Cp), dtype=torch.float32, device=W1.device)
        # layer 1: propagation + bias + (act) + dropout, layer 2's thin projection fused in the epilogue
        hit = cache is not None and cache.h1 is not None and cache.key == cache_key
        fuse = C <= ops.FUSED_PROJ_MAX_CLASSES and not hit
        dkw = dict(drop_mode=drop.mode, drop_p=drop.p, keep_mask=drop.mask, philox_seed=drop.seed, philox_offset=drop.offset)
        if hit:
            H1d = cache.h1 if (drop.mode == ops.DROP_NONE or drop.p <= 0.0) else ops.dropout_apply(cache.h1, F=Hp, **dkw)
        else:
            B1 = _identity_operand(W1, feat)
            H1d, _ = ops.spmm(graph, B1, F=Hp, bias=b1, act=act, W_proj=W2p if fuse else None, P=P if fuse else None, **dkw)
            if cache is not None and (drop.mode == ops.DROP_NONE or drop.p <= 0.0):
                cache.key, cache.h1 = cache_key, H1d          # pre-dropout activation of this parameter state
        if not fuse:
            ops.project(H1d, W2p, K=Hp, out=P)
        # layer 2: propagation of the projected rows + bias
        Z2, _ = ops.spmm(graph, P, F=Cp, bias=b2)
        ctx.graph, ctx.feat, ctx.act, ctx.drop = graph, feat, act, drop
        ctx.dims = (n, H, C, Hp, Cp, int(W1.shape[0]))
        ctx.save_for_backward(H1d, W2p)
        return Z2[:, :C] if Cp != C else Z2

    @staticmethod
    def backward(ctx, dlogits):
        graph, feat, act, drop = ctx.graph, ctx.feat, ctx.act, ctx.drop
        n, H, C, Hp, Cp, in_ch = ctx.dims
        H1d, W2p = ctx.saved_tensors
        gt = graph.transpose()
        if Cp == C and dlogits.is_contiguous() and dlogits.data_ptr() % 16 == 0:
            dZ2 = dlogits
        else:
            dZ2 = torch.zeros((n, Cp), dtype=torch.float32, device=dlogits.device)
            dZ2[:, :C].copy_(dlogits)
        G2, _ = ops.spmm(gt, dZ2, F=Cp)                                  # A^T dZ2
        r = ops.dense_bwd(G2, H1d, W2p, dZ2, H=Hp, n_classes=C, act=act, drop_mode=drop.mode, drop_p=drop.p,
                          keep_mask=drop.mask, philox_seed=drop.seed, philox_offset=drop.offset)
        dW1 = torch.empty((in_ch, H), dtype=torch.float32, device=dlogits.device)
        if Hp == H:
            ops.spmm(gt, r["dZ1"], F=Hp, out=dW1)                        # dW1[:N] = A^T dZ1  (X = I)
            G1 = dW1
        else:
            G1, _ = ops.spmm(gt, r["dZ1"], F=Hp)
            dW1[:n].copy_(G1[:, :H])
        if feat.Fdoc is not None:
            tail = torch.empty((in_ch - n, H), dtype=torch.float32, device=dlogits.device)
            if Hp == H:
                ops.hier_backward(G1, n, feat.n_vocab, feat.Fdoc, H, tail)
            else:
                ops.hier_backward(G1[:, :H].contiguous(), n, feat.n_vocab, feat.Fdoc, H, tail)
            dW1[n:].copy_(tail)
        return dW1, r["db_hidden"][:H], r["dW2"][:H], r["db_out"], None, None, None, None, None, None


# --------------------------------------------------------------------------------------
# generic single layer (dense x, or [I | F] x):  out = A (x W) + b
# --------------------------------------------------------------------------------------
class _GCNConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, graph: GraphCSR, feat: Optional[_FeatInfo]):
        n = graph.n_nodes
        out_ch = int(W.shape[1])
        Fp = ops.pad4(out_ch)
        if feat is not None:
            XW = _identity_operand(W, feat)
        else:
            XW = torch.zeros((n, Fp), dtype=torch.float32, device=W.device) if Fp != out_ch else \
                torch.empty((n, Fp), dtype=torch.float32, device=W.device)
            if out_ch <= 256 and W.shape[0] <= 512:
                ops.project(x.contiguous(), W.contiguous(), out=XW)       # thin: hand-written row kernel
            else:
                torch.matmul(x, W, out=XW[:, :out_ch]) if Fp == out_ch else XW[:, :out_ch].copy_(torch.matmul(x, W))
        out, _ = ops.spmm(graph, XW, F=Fp, bias=b)
        ctx.graph, ctx.feat = graph, feat
        ctx.dims = (n, out_ch, Fp, int(W.shape[0]))
        ctx.save_for_backward(x if feat is None else None, W)
        return out[:, :out_ch] if Fp != out_ch else out

    @staticmethod
    def backward(ctx, dout):
        graph, feat = ctx.graph, ctx.feat
        n, out_ch, Fp, in_ch = ctx.dims
        x, W = ctx.saved_tensors
        gt = graph.transpose()
        if Fp == out_ch and dout.is_contiguous() and dout.data_ptr() % 16 == 0:
            d = dout
        else:
            d = torch.zeros((n, Fp), dtype=torch.float32, device=dout.device)
            d[:, :out_ch].copy_(dout)
        G, _ = ops.spmm(gt, d, F=Fp)            # gradient wrt (x W)
        G = G[:, :out_ch]
        db = dout.sum(dim=0)
        if feat is not None:
            dW = torch.empty((in_ch, out_ch), dtype=torch.float32, device=dout.device)
            dW[:n].copy_(G)
            if feat.Fdoc is not None:
                tail = torch.empty((in_ch - n, out_ch), dtype=torch.float32, device=dout.device)
                ops.hier_backward(G.contiguous(), n, feat.n_vocab, feat.Fdoc, out_ch, tail)
                dW[n:].copy_(tail)
            return None, dW, db, None, None
        dW = x.t().matmul(G)                     # plain library GEMMs on the generic dense path
        dx = G.matmul(W.t()) if ctx.needs_input_grad[0] else None
        return dx, dW, db, None, None


class GCNConv(nn.Module):
    """`torch_geometric.nn.GCNConv(in, out, add_self_loops=True)` look-alike
    (reference call sites: textgcn/lib/models.py:11-15,20).  normalize=True, cached graph
    upload (values identical to recomputing gcn_norm each call), bias=True."""

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if improved or not add_self_loops or not normalize:
            raise NotImplementedError("pytextgcn_b200.GCNConv implements the configuration the reference uses: "
                                      "improved=False, add_self_loops=True, normalize=True")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _glorot_(self.weight.data)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor] = None,
                n_vocab: Optional[int] = None, holder=None) -> torch.Tensor:
        if not self.weight.is_cuda:
            raise RuntimeError("pytextgcn_b200.GCNConv runs on CUDA only: move the module and the graph to a "
                               "CUDA device (gcn.to('cuda'); g.to('cuda')) -- there is no CPU path")
        n = int(x.shape[0])
        graph = get_graph(edge_index, edge_weight, n, holder=holder)
        feat = decode_features(x, n_vocab) if x.is_sparse else None
        if x.is_sparse and feat is None:
            x = x.to_dense()           # general sparse features: not a TextGCN input; dense path
        bias = self.bias if self.bias is not None else torch.zeros(self.out_channels, device=self.weight.device)
        return _GCNConvFunction.apply(None if feat is not None else x, self.weight, bias, graph, feat)

    def __repr__(self) -> str:
        return f"GCNConv({self.in_channels}, {self.out_channels})"


class GCN(nn.Module):
    """Drop-in for `textgcn.lib.models.GCN` (textgcn/lib/models.py:6-25)."""

    def __init__(self, in_channels, out_channels, n_gcn=2, n_hidden_gcn=64, activation=nn.ReLU, dropout=0.5,
                 apply_activation: bool = False):
        super().__init__()
        self.activation = activation()         # constructed, unused -- as in the reference (models.py:9,22)
        self.dropout = dropout
        self.apply_activation = apply_activation
        if apply_activation and not isinstance(self.activation, nn.ReLU):
            raise NotImplementedError("the fused epilogue implements ReLU only")
        self.layers = nn.ModuleList([GCNConv(in_channels, n_hidden_gcn, add_self_loops=True)])
        for _ in range(n_gcn - 2):
            self.layers.append(GCNConv(n_hidden_gcn, n_hidden_gcn, add_self_loops=True))
        self.layers.append(GCNConv(n_hidden_gcn, out_channels, add_self_loops=True))
        self._drop_calls = 0
        self.drop_mask_override = None    # list of bool keep-masks, one per hidden layer (parity tests)
        self.seed = None                  # Philox key; drawn from torch's generator on first use
        self.share_hidden = True          # reuse the eval forward's hidden activation in the next training forward
        self._hidden_cache = _HiddenCache()

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_hidden_cache"] = _HiddenCache()       # th.save(gcn) (flat_amazon.py:128) must not pickle a cached activation
        return d

    def invalidate_cache(self) -> None:
        """Forget the cached hidden activation (needed only when W1/b1 are changed behind torch's back, e.g. through
        raw device pointers; in-place torch ops such as optimizer.step() are seen through the version counters)."""
        self._hidden_cache = _HiddenCache()

    def _dropout_spec(self, layer_idx: int) -> DropoutSpec:
        if not self.training or self.dropout <= 0.0:
            return _NO_DROP
        if self.drop_mask_override is not None:
            m = self.drop_mask_override[layer_idx]
            return DropoutSpec(ops.DROP_MASK, self.dropout, m.to(torch.uint8) if m.dtype == torch.bool else m)
        if self.seed is None:
            self.seed = int(torch.randint(0, 2**62, (1,)).item())
        self._drop_calls += 1
        return DropoutSpec(ops.DROP_PHILOX, self.dropout, None, self.seed, self._drop_calls)

    def forward(self, g) -> torch.Tensor:
        x = g.x
        if not self.layers[0].weight.is_cuda:
            raise RuntimeError("pytextgcn_b200.GCN runs on CUDA only (gcn.to('cuda'); g.to('cuda')); "
                               "there is no CPU path")
        n_vocab = getattr(g, "n_vocab", None)
        act = ops.ACT_RELU if self.apply_activation else ops.ACT_NONE
        feat = decode_features(x, n_vocab) if x.is_sparse else None
        if len(self.layers) == 2 and feat is not None:
            graph = get_graph(g.edge_index, g.edge_attr, int(x.shape[0]), holder=g)
            l0, l1 = self.layers[0], self.layers[1]
            cache, key = None, None
            if self.share_hidden:
                cache = getattr(self, "_hidden_cache", None)
                if cache is None:                      # modules unpickled from an older checkpoint
                    cache = self._hidden_cache = _HiddenCache()
                key = (l0.weight._version, l0.bias._version, l0.weight.data_ptr(), l0.bias.data_ptr(), id(graph), act)
            return _GCN2Function.apply(l0.weight, l0.bias, l1.weight, l1.bias, graph, feat, act, self._dropout_spec(0),
                                       cache, key)
        # generic depth / dense features: layer by layer (activation/dropout as torch elementwise ops)
        for i, layer in enumerate(self.layers):
            x = layer(x, g.edge_index, g.edge_attr, n_vocab=n_vocab, holder=g)
            if i < len(self.layers) - 1:
                if self.apply_activation:
                    x = torch.relu(x)
                if self.training and self.dropout > 0:
                    if self.drop_mask_override is not None:
                        x = x * self.drop_mask_override[i].to(x.dtype) * (1.0 / (1.0 - self.dropout))
                    else:
                        x = nn.functional.dropout(x, p=self.dropout, training=True)
        return x
