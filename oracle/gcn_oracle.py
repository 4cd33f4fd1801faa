"""CPU oracle for the TextGCN training hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-torch (CPU, fp32 by default) restatement of the arithmetic the
reference executes when `textgcn/lib/models.py:17-25` (`GCN.forward`) drives
`torch_geometric.nn.GCNConv` and `flat_amazon.py:82,101-106` computes the masked
cross-entropy, back-propagates and steps Adam.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product package (`pytextgcn_b200/`) never does.

PARITY UNPINNED at the GCNConv boundary: the arithmetic lives in the third-party package
`torch-geometric==1.6.3` (+ `torch-scatter==2.0.5`, `pytorch=1.7.0`;
`/root/reference/requirements.yml:44,87-88`), which is neither vendored under
`/root/reference` nor installable here (no wheel, no network), and the reference's own
tests hold no golden logits/gradients for this path (`textgcn/test/test_model.py:10-40`
asserts nothing).  What anchors this restatement instead:
  * the published algorithm (Kipf & Welling 2017, eq. 2/9: A_hat = D^-1/2 (A+I) D^-1/2),
    restated a second time, independently, as a dense fp64 matrix product
    (`dense_ahat_fp64`, `dense_forward_fp64`) and cross-checked in tests/test_oracle.py;
  * the reference call sites: `GCNConv(in, out, add_self_loops=True)` with defaults
    (normalize=True, cached=False, improved=False, bias=True) at `models.py:11-15`,
    `layer(x, g.edge_index, g.edge_attr)` at `models.py:20`, dropout-but-no-activation at
    `models.py:21-23`;
  * torch autograd on this restatement (gradcheck in fp64) for the backward formulas.

PyG-1.6.3 semantics restated here (recalled from upstream `gcn_conv.py`, `utils/loop.py`,
`message_passing.py`):
  add_remaining_self_loops: drop edges with row==col, append N loops (weight 1.0, or the
      weight of a pre-existing loop on that node) AFTER the original edges;
  gcn_norm: deg = scatter_add(w, col=edge_index[1]); dis = deg.pow_(-0.5); inf -> 0;
      w_hat = dis[row] * w * dis[col]   (left to right);
  GCNConv.forward: xw = x @ weight (weight is (in, out); x may be sparse COO);
      out[i] = sum_{e: col_e = i} w_hat_e * xw[row_e];  out += bias.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# gcn_norm  [PyG-1.6.3 torch_geometric/nn/conv/gcn_conv.py::gcn_norm, utils/loop.py]
# --------------------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: Tensor, edge_weight: Tensor, fill_value: float,
                             num_nodes: int) -> Tuple[Tensor, Tensor]:
    """[PyG-1.6.3 utils/loop.py::add_remaining_self_loops] -- call site models.py:20."""
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = torch.arange(0, num_nodes, dtype=row.dtype, device=row.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    inv_mask = ~mask
    loop_weight = torch.full((num_nodes,), fill_value, dtype=edge_weight.dtype,
                             device=edge_weight.device)
    remaining = edge_weight[inv_mask]
    if remaining.numel() > 0:
        loop_weight[row[inv_mask]] = remaining
    edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, edge_weight


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int,
             dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """[PyG-1.6.3 gcn_conv.py::gcn_norm, improved=False, add_self_loops=True].

    Returns (edge_index', w_hat) with the N self loops appended after the original edges.
    On CPU, `scatter_add_` over a 1-D index is a sequential fp32 sum in edge order and
    `pow_(-0.5)` is bit-identical to IEEE 1.0f/sqrtf(x) (both verified).
    """
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, 1.0, num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).scatter_add_(0, col, edge_weight)
    deg_inv_sqrt = deg.pow_(-0.5)
    deg_inv_sqrt.masked_fill_(deg_inv_sqrt == float("inf"), 0)
    w_hat = deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]
    return edge_index, w_hat


def gcn_norm_with_dis(edge_index: Tensor, edge_weight: Tensor, num_nodes: int):
    """Same as gcn_norm but also returns deg^-1/2 (needed to check the device `dis`)."""
    ei, w = add_remaining_self_loops(edge_index, edge_weight, 1.0, num_nodes)
    row, col = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=w.dtype).scatter_add_(0, col, w)
    dis = deg.pow_(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return ei, dis[row] * w * dis[col], dis


def csr_from_gcn_norm(edge_index: Tensor, edge_weight: Tensor, num_nodes: int):
    """CSR of A_hat keyed by TARGET node (col = edge_index[1], the scatter side of
    propagate), built by a STABLE sort of gcn_norm's output, so every CSR row lists its
    in-edges in original edge order followed by its self loop.

    Returns rowptr int64[N+1], colidx int64[nnz] (the SOURCE node of each entry), val
    fp32[nnz], dis fp32[N], perm int64[nnz] (position in gcn_norm's edge list).
    This is what `tgcn_csr_from_coo_gcn_norm` must reproduce bit for bit.
    """
    ei, w_hat, dis = gcn_norm_with_dis(edge_index, edge_weight, num_nodes)
    src, dst = ei[0], ei[1]
    perm = torch.sort(dst, stable=True).indices
    counts = torch.bincount(dst, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, src[perm].contiguous(), w_hat[perm].contiguous(), dis, perm


# --------------------------------------------------------------------------------------
# GCNConv / GCN forward  [PyG-1.6.3 GCNConv.forward; textgcn/lib/models.py:17-25]
# --------------------------------------------------------------------------------------
def gcn_conv(x: Tensor, edge_index: Tensor, edge_weight: Tensor, weight: Tensor,
             bias: Optional[Tensor]) -> Tensor:
    """One `GCNConv(in, out, add_self_loops=True)` call (cached=False: gcn_norm recomputed
    on every call, as the reference does -- models.py:11-15,20)."""
    n = x.size(0)
    ei, w_hat = gcn_norm(edge_index, edge_weight, n, dtype=weight.dtype)
    if w_hat.dtype != weight.dtype:
        w_hat = w_hat.to(weight.dtype)
    xw = torch.matmul(x, weight)                 # sparse COO x dense for layer 1
    x_j = xw.index_select(0, ei[0])              # gather on source
    msg = w_hat.view(-1, 1) * x_j                # message
    out = torch.zeros(n, xw.size(1), dtype=xw.dtype).index_add_(0, ei[1], msg)  # aggr='add' on target
    if bias is not None:
        out = out + bias
    return out


def gcn_forward(x: Tensor, edge_index: Tensor, edge_attr: Tensor,
                weights: Sequence[Tensor], biases: Sequence[Optional[Tensor]],
                p: float = 0.5, training: bool = False,
                drop_masks: Optional[Sequence[Tensor]] = None,
                relu: bool = False) -> Tensor:
    """`GCN.forward` (models.py:17-25): conv, then dropout after every layer but the last;
    NO activation (models.py:22 is commented out).  `relu=True` inserts the TextGCN-paper
    activation before dropout -- not reference behaviour, used only to check the optional
    fused-ReLU epilogue.  `drop_masks[i]` (bool keep-mask, N x hidden) replaces the RNG so
    the CUDA path can be compared element for element; scaling is 1/(1-p) as F.dropout.
    """
    h = x
    n_layers = len(weights)
    for i in range(n_layers):
        h = gcn_conv(h, edge_index, edge_attr, weights[i], biases[i])
        if i < n_layers - 1:
            if relu:
                h = torch.relu(h)
            if training and p > 0.0:
                if drop_masks is not None:
                    h = h * drop_masks[i].to(h.dtype) * (1.0 / (1.0 - p))
                else:
                    h = torch.nn.functional.dropout(h, p=p, training=True)
    return h


def masked_cross_entropy(logits: Tensor, y: Tensor, mask: Tensor) -> Tensor:
    """`CrossEntropyLoss(reduction='mean')(gcn(g)[mask], g.y[mask])` -- flat_amazon.py:82,101-102."""
    return torch.nn.functional.cross_entropy(logits[mask], y[mask], reduction="mean")


def sparse_identity_features(n_nodes: int, hierarchy_feats: Optional[Tensor] = None,
                             n_vocab: int = 0) -> Tensor:
    """X = I_N (sparse COO) or [I_N | F] with F on the document rows only --
    text2graph.py:226-246."""
    idx = torch.arange(n_nodes, dtype=torch.int64)
    inds = torch.stack([idx, idx])
    vals = torch.ones(n_nodes, dtype=torch.float32)
    n_cols = n_nodes
    if hierarchy_feats is not None:
        hf = hierarchy_feats.to(torch.float32)
        r, c = torch.nonzero(hf, as_tuple=True)
        inds = torch.cat([inds, torch.stack([r + n_vocab, c + n_nodes])], dim=1)
        vals = torch.cat([vals, hf[r, c]])
        n_cols = n_nodes + hf.shape[1]
    return torch.sparse_coo_tensor(inds, vals, size=(n_nodes, n_cols), dtype=torch.float32).coalesce()


def glorot_(weight: Tensor, generator: Optional[torch.Generator] = None) -> Tensor:
    """[PyG-1.6.3 nn/inits.py::glorot]  U(-a, a), a = sqrt(6 / (in + out))."""
    a = math.sqrt(6.0 / (weight.size(-2) + weight.size(-1)))
    with torch.no_grad():
        weight.uniform_(-a, a, generator=generator)
    return weight


# --------------------------------------------------------------------------------------
# Independent second formulation: dense fp64  (Kipf & Welling eq. 2)
# --------------------------------------------------------------------------------------
def dense_ahat_fp64(edge_index: Tensor, edge_weight: Tensor, num_nodes: int) -> Tensor:
    """A_hat[i, j] = weight of edge j -> i, normalised; A_hat = D^-1/2 (A + I) D^-1/2 with
    D = in-degree (row sums of A + I in this orientation).  Small graphs only."""
    A = torch.zeros(num_nodes, num_nodes, dtype=torch.float64)
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    A.index_put_((dst[keep], src[keep]), edge_weight[keep].to(torch.float64), accumulate=True)
    diag = torch.ones(num_nodes, dtype=torch.float64)
    if (~keep).any():
        diag[src[~keep]] = edge_weight[~keep].to(torch.float64)
    A = A + torch.diag(diag)
    deg = A.sum(dim=1)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    # value on edge j->i is dis[j] * w * dis[i]
    return dis.view(-1, 1) * A * dis.view(1, -1)


def dense_forward_fp64(x_dense: Tensor, edge_index: Tensor, edge_attr: Tensor,
                       weights: Sequence[Tensor], biases: Sequence[Tensor],
                       p: float = 0.0, drop_masks: Optional[Sequence[Tensor]] = None,
                       relu: bool = False) -> Tensor:
    n = x_dense.size(0)
    ahat = dense_ahat_fp64(edge_index, edge_attr, n)
    h = x_dense.to(torch.float64)
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = ahat @ (h @ w.to(torch.float64)) + b.to(torch.float64)
        if i < len(weights) - 1:
            if relu:
                h = torch.relu(h)
            if drop_masks is not None and p > 0:
                h = h * drop_masks[i].to(torch.float64) / (1.0 - p)
    return h


# --------------------------------------------------------------------------------------
# Reference epoch (flat_amazon.py:99-117) -- used for gradients-of-record and CPU timing
# --------------------------------------------------------------------------------------
class OracleGCN(torch.nn.Module):
    """Same parameter names/layout as the reference module: layers.{i}.weight (in,out),
    layers.{i}.bias (out)  (models.py:11-15 + PyG GCNConv)."""

    class _Layer(torch.nn.Module):
        def __init__(self, cin, cout):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.empty(cin, cout))
            self.bias = torch.nn.Parameter(torch.zeros(cout))
            glorot_(self.weight)

    def __init__(self, in_channels, out_channels, n_gcn=2, n_hidden_gcn=64, dropout=0.5, relu=False):
        super().__init__()
        dims = [in_channels] + [n_hidden_gcn] * (n_gcn - 1) + [out_channels]
        self.layers = torch.nn.ModuleList([self._Layer(dims[i], dims[i + 1]) for i in range(n_gcn)])
        self.dropout = dropout
        self.relu = relu

    def forward(self, g, drop_masks=None):
        return gcn_forward(g.x, g.edge_index, g.edge_attr,
                           [l.weight for l in self.layers], [l.bias for l in self.layers],
                           p=self.dropout, training=self.training, drop_masks=drop_masks,
                           relu=self.relu)


def reference_epoch(gcn: OracleGCN, g, optimizer, drop_masks=None):
    """One reference 'epoch' = train step + eval forward + val loss + host argmax
    (flat_amazon.py:99-116; sklearn f1/accuracy replaced by a numpy accuracy)."""
    gcn.train()
    outputs = gcn(g, drop_masks=drop_masks)[g.train_mask]
    loss = torch.nn.functional.cross_entropy(outputs, g.y[g.train_mask], reduction="mean")
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    gcn.eval()
    with torch.no_grad():
        logits = gcn(g)
        val_loss = torch.nn.functional.cross_entropy(logits[g.val_mask], g.y[g.val_mask],
                                                     reduction="mean")
        pred_val = logits[g.val_mask].cpu().numpy().argmax(axis=1)
        pred_train = logits[g.train_mask].cpu().numpy().argmax(axis=1)
        acc_val = float((pred_val == g.y[g.val_mask].numpy()).mean()) if pred_val.size else 0.0
        acc_train = float((pred_train == g.y[g.train_mask].numpy()).mean()) if pred_train.size else 0.0
    return loss.item(), val_loss.item(), acc_train, acc_val
