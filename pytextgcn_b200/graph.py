"""Graph upload: `Data.edge_index / edge_attr` -> device CSR of A_hat, once per graph.

The reference recomputes gcn_norm inside every GCNConv call (cached=False,
textgcn/lib/models.py:11-15,20) -- four times per epoch.  Here the COO graph is converted once
into a CSR keyed by target node whose values are bit-identical to gcn_norm's (see
csrc/csr_build.cu), plus the row-chunk plan the SpMM kernels consume.  The result is cached on
the `Data` object / per edge_index storage so the per-epoch calls do no preprocessing.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import _native

DEFAULT_CHUNK_NNZ = None      # None -> auto_chunk_nnz()


def auto_chunk_nnz(nnz: int) -> int:
    """Chunk length for the SpMM plan: about nnz/5000 rounded to a power of two in [256, 2048].
    Long chunks minimise split rows (partial-row scratch traffic), but one warp walks a chunk
    serially (~8 gathers in flight), so the longest chunk must stay a small fraction of the
    whole launch -- measured on B200: 2048 is best at 2.2e7 nnz, 512 at 2.7e6 (a 1/8 shard)."""
    import math
    c = max(nnz, 1) / 5000.0
    return int(min(2048, max(256, 2 ** round(math.log2(max(c, 1.0))))))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class SpmmPlan:
    """Row-chunk work list for a row range of the CSR (tgcn_spmm_plan)."""
    row_begin: int
    row_end: int
    chunk_nnz: int
    chunks: torch.Tensor          # int32 [n_chunks, 4] = {row, begin, end, slot}
    n_chunks: int
    split_rows: torch.Tensor      # int32 [n_split_rows, 3] = {row, first_slot, n_slots}
    n_split_rows: int
    n_slots: int
    max_row_nnz: int
    slot_owner: Optional[torch.Tensor] = None    # int32 [n_slots]: split-row index of each scratch slot
    counters: Optional[torch.Tensor] = None      # int32 [n_split_rows], zero; arrival counters (self-resetting)
    _scratch: Dict[int, torch.Tensor] = field(default_factory=dict)

    def scratch(self, F: int) -> Optional[torch.Tensor]:
        if self.n_slots == 0:
            return None
        t = self._scratch.get(F)
        if t is None:
            t = torch.empty((self.n_slots, F), dtype=torch.float32, device=self.chunks.device)
            self._scratch[F] = t
        return t


class GraphCSR:
    """Device-resident CSR of A_hat = D^-1/2 (A+I) D^-1/2, rows = target nodes."""

    def __init__(self, n_nodes: int, rowptr: torch.Tensor, colidx: torch.Tensor, val: torch.Tensor,
                 dis: Optional[torch.Tensor], edge_slot: Optional[torch.Tensor] = None, n_cols: Optional[int] = None):
        self.n_nodes = n_nodes                     # rows of this CSR
        self.n_cols = n_nodes if n_cols is None else n_cols   # > n_nodes for a row shard (1D partition)
        self.rowptr = rowptr
        self.colidx = colidx
        self.val = val
        self.dis = dis
        self.edge_slot = edge_slot
        self.nnz = int(colidx.numel())
        self.device = rowptr.device
        self._plans: Dict[Tuple[int, int, int], SpmmPlan] = {}
        self._symmetric: Optional[bool] = None
        self._transpose: Optional["GraphCSR"] = None
        self.buffers: Dict[str, torch.Tensor] = {}   # static work buffers owned by the host layer

    # ---- SpMM plan ----
    def plan(self, row_begin: int = 0, row_end: Optional[int] = None,
             chunk_nnz: Optional[int] = DEFAULT_CHUNK_NNZ, sort_chunks: bool = True) -> SpmmPlan:
        row_end = self.n_nodes if row_end is None else row_end
        if chunk_nnz is None:
            chunk_nnz = auto_chunk_nnz(self.nnz if (row_begin == 0 and row_end == self.n_nodes) else
                                       int(self.rowptr[row_end].item()) - int(self.rowptr[row_begin].item()))
        key = (row_begin, row_end, chunk_nnz, sort_chunks)
        p = self._plans.get(key)
        if p is not None:
            return p
        lib = _native.load()
        n_rows = row_end - row_begin
        with torch.cuda.device(self.device):
            nnz_range = int(self.rowptr[row_end].item()) - int(self.rowptr[row_begin].item())
            cap = n_rows + nnz_range // chunk_nnz + 1
            chunks = torch.empty((cap, 4), dtype=torch.int32, device=self.device)
            split = torch.empty((max(nnz_range // chunk_nnz + 1, 1), 3), dtype=torch.int32, device=self.device)
            counts = torch.zeros(4, dtype=torch.int32, device=self.device)
            owner = torch.zeros(cap, dtype=torch.int32, device=self.device)
            _native.check(lib.tgcn_spmm_plan(self.rowptr.data_ptr(), row_begin, row_end, chunk_nnz,
                                             chunks.data_ptr(), cap, split.data_ptr(), owner.data_ptr(), counts.data_ptr(),
                                             None, 0, _stream()))
            n_chunks, n_slots, n_split, max_len = (int(v) for v in counts.cpu().tolist())
        chunks = chunks[:n_chunks]
        if sort_chunks and n_chunks > 1:
            # longest chunks first, neighbours of similar length: the 8 warps of a CTA finish together
            # (no idle warps holding an SM slot) and the tail of the grid is made of short rows
            lens = chunks[:, 2] - chunks[:, 1]
            chunks = chunks[torch.argsort(lens, descending=True, stable=True)].contiguous()
        p = SpmmPlan(row_begin, row_end, chunk_nnz, chunks, n_chunks, split[:max(n_split, 0)],
                     n_split, n_slots, max_len,
                     slot_owner=owner[:max(n_slots, 1)].clone() if n_split else None,
                     counters=torch.zeros(max(n_split, 1), dtype=torch.int32, device=self.device) if n_split else None)
        self._plans[key] = p
        return p

    # ---- restrictions to the rows / columns a masked loss can reach ----
    def plan_for_rows(self, row_mask: torch.Tensor, base: Optional[SpmmPlan] = None) -> SpmmPlan:
        """The chunks of `base` (default: the whole-matrix plan) whose row is selected by the boolean `row_mask`:
        the SpMM then computes exactly those output rows and leaves the others untouched.  Split rows keep
        their scratch slots and arrival counters (a row is in or out as a whole)."""
        base = base or self.plan()
        keep = row_mask.to(self.device)[base.chunks[:, 0].to(torch.int64) - 0]
        chunks = base.chunks[keep].contiguous()
        return SpmmPlan(base.row_begin, base.row_end, base.chunk_nnz, chunks, int(chunks.shape[0]), base.split_rows,
                        base.n_split_rows, base.n_slots, base.max_row_nnz, slot_owner=base.slot_owner,
                        counters=base.counters, _scratch=base._scratch)

    def select_columns(self, col_mask: torch.Tensor) -> "GraphCSR":
        """CSR of the same rows with only the entries whose COLUMN is selected (entry order kept).  For an
        operand that is zero outside the selected rows -- the loss gradient dZ2 outside the training rows --
        A_hat B and this matrix times B are the same sums without the zero terms."""
        keep = col_mask.to(self.device)[self.colidx.to(torch.int64)]
        rows = self.row_ids()[keep]
        counts = torch.bincount(rows, minlength=self.n_nodes)
        rowptr = torch.zeros(self.n_nodes + 1, dtype=torch.int32, device=self.device)
        rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        sub = GraphCSR(self.n_nodes, rowptr, self.colidx[keep].contiguous(), self.val[keep].contiguous(), self.dis,
                       n_cols=self.n_cols)
        sub._symmetric = False
        return sub

    def plan_nonempty(self, chunk_nnz: Optional[int] = DEFAULT_CHUNK_NNZ) -> SpmmPlan:
        """Plan over the rows that hold at least one entry (rows without entries are never written)."""
        nonempty = (self.rowptr[1:] - self.rowptr[:-1]) > 0
        return self.plan_for_rows(nonempty, self.plan(chunk_nnz=chunk_nnz))

    # ---- transpose handling for the backward pass ----
    def row_ids(self) -> torch.Tensor:
        counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        return torch.repeat_interleave(torch.arange(self.n_nodes, device=self.device), counts)

    def is_symmetric(self) -> bool:
        """A_hat == A_hat^T (true for Text2GraphTransformer graphs: PMI and TF-IDF edges are
        emitted in both directions with equal weights, text2graph.py:148-170).  Checked once, on
        the device, by comparing the sorted (row, col, value) triples of A_hat and A_hat^T (values to
        within 1e-6 relative: the two multiplication orders may round differently in the last bit)."""
        if self._symmetric is None:
            rows = self.row_ids()
            cols = self.colidx.to(torch.int64)
            n = self.n_nodes
            k1 = rows * n + cols
            k2 = cols * n + rows
            o1 = torch.argsort(k1)
            o2 = torch.argsort(k2)
            same_pattern = bool(torch.equal(k1[o1], k2[o2]))
            same_vals = False
            if same_pattern:
                # (dis[j]*w)*dis[i] and (dis[i]*w)*dis[j] may differ in the last bit: compare to a few ulp
                a, b = self.val[o1], self.val[o2]
                same_vals = bool(((a - b).abs() <= 1e-6 * torch.maximum(a.abs(), b.abs())).all())
            self._symmetric = same_vals
        return self._symmetric

    def transpose(self) -> "GraphCSR":
        """CSR of A_hat^T (only built for non-symmetric graphs; one-off, torch index ops)."""
        if self.is_symmetric():
            return self
        if self._transpose is None:
            rows = self.row_ids()
            cols = self.colidx.to(torch.int64)
            perm = torch.sort(cols, stable=True).indices
            counts = torch.bincount(cols, minlength=self.n_nodes)
            rowptr = torch.zeros(self.n_nodes + 1, dtype=torch.int32, device=self.device)
            rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            t = GraphCSR(self.n_nodes, rowptr, rows[perm].to(torch.int32).contiguous(), self.val[perm].contiguous(), self.dis)
            t._symmetric = False
            t._transpose = self
            self._transpose = t
        return self._transpose

    def buffer(self, name: str, shape, dtype=torch.float32, zero: bool = False) -> torch.Tensor:
        """Named static work buffer (stable address across steps -> CUDA-graph friendly)."""
        t = self.buffers.get(name)
        shape = tuple(shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self.buffers[name] = t
        return t

    def csr_bytes(self) -> int:
        return self.nnz * 8 + (self.n_nodes + 1) * 4


def upload_graph(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], num_nodes: int,
                 keep_edge_slot: bool = False) -> GraphCSR:
    """COO -> CSR of A_hat on the device of `edge_index` (must be CUDA).  One host sync.

    edge_index: int64 (2, E), any strides (the reference emits the transposed view `coo.T`,
    text2graph.py:171,192 -- consumed in place).  edge_attr: fp32 (E) or None (all ones).
    """
    if not edge_index.is_cuda:
        raise RuntimeError("pytextgcn_b200.upload_graph: edge_index must live on a CUDA device "
                           "(there is no CPU path; move the Data object with g.to('cuda'))")
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise RuntimeError("edge_index must be an int64 tensor of shape (2, E)")
    lib = _native.load()
    dev = edge_index.device
    E = int(edge_index.size(1))
    N = int(num_nodes)
    if edge_attr is not None:
        if edge_attr.numel() != E:
            raise RuntimeError(f"edge_attr has {edge_attr.numel()} entries, edge_index has {E} edges")
        edge_attr = edge_attr.to(device=dev, dtype=torch.float32).contiguous().view(-1)
    s0, s1 = edge_index.stride(0), edge_index.stride(1)
    if E > 0 and s1 < 1:
        edge_index = edge_index.contiguous()
        s0, s1 = edge_index.stride(0), edge_index.stride(1)
    base = edge_index.data_ptr()
    src_ptr, dst_ptr = base, base + s0 * 8
    with torch.cuda.device(dev):
        ws_bytes = C.c_size_t(0)
        _native.check(lib.tgcn_csr_workspace_bytes(N, E, C.byref(ws_bytes)))
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
        rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        colidx = torch.empty(E + N, dtype=torch.int32, device=dev)
        val = torch.empty(E + N, dtype=torch.float32, device=dev)
        dis = torch.empty(N, dtype=torch.float32, device=dev)
        slot = torch.empty(E + N, dtype=torch.int32, device=dev) if keep_edge_slot else None
        status = torch.zeros(2, dtype=torch.int32, device=dev)
        _native.check(lib.tgcn_csr_from_coo_gcn_norm(
            src_ptr, dst_ptr, max(s1, 1), _native.ptr(edge_attr), E, N,
            rowptr.data_ptr(), colidx.data_ptr(), val.data_ptr(), dis.data_ptr(),
            _native.ptr(slot), status.data_ptr(), ws.data_ptr(), ws_bytes.value, _stream()))
        st, nnz = (int(v) for v in status.cpu().tolist())   # host sync
        del ws
    if st != 0:
        raise RuntimeError(f"edge_index holds node ids outside [0, {N}) (textgcn_b200 error {st})")
    if nnz < E + N:   # self loops were dropped: shrink (copy so the big buffers are released)
        colidx = colidx[:nnz].clone()
        val = val[:nnz].clone()
    return GraphCSR(N, rowptr, colidx, val, dis, slot)


# ---- CSR sidecar on disk: lets a later run skip the sort of the upload -------------
def _fingerprint(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], n: int):
    """Cheap content check tying a sidecar to the graph it was built from (exact integer / fp64 sums)."""
    e = int(edge_index.shape[1])
    s_idx = int((edge_index[0].to(torch.int64) * 31 + edge_index[1].to(torch.int64)).sum().item()) if e else 0
    s_w = float(edge_attr.double().sum().item()) if (edge_attr is not None and e) else 0.0
    return [n, e, s_idx, s_w]


def save_csr(graph: GraphCSR, path: str, edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor]) -> None:
    """Writes the device CSR next to a pickled `Data` (text2graph.py:195-202 writes TGData_<time>.p)."""
    torch.save({"format": "textgcn_b200.csr.v1", "n_nodes": graph.n_nodes, "fingerprint": _fingerprint(edge_index, edge_attr, graph.n_nodes),
                "rowptr": graph.rowptr.cpu(), "colidx": graph.colidx.cpu(), "val": graph.val.cpu(), "dis": graph.dis.cpu()}, path)


def load_csr(path: str, edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], num_nodes: int) -> Optional[GraphCSR]:
    """Returns the CSR stored at `path` if it belongs to this graph (fingerprint match), else None."""
    import os
    if not os.path.exists(path):
        return None
    try:
        d = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        return None
    if d.get("format") != "textgcn_b200.csr.v1" or d.get("fingerprint") != _fingerprint(edge_index, edge_attr, num_nodes):
        return None
    dev = edge_index.device
    return GraphCSR(int(d["n_nodes"]), d["rowptr"].to(dev), d["colidx"].to(dev), d["val"].to(dev), d["dis"].to(dev))


def upload_graph_cached(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], num_nodes: int, sidecar: str) -> GraphCSR:
    """upload_graph with a disk sidecar: load when it matches the graph, otherwise build and write it."""
    g = load_csr(sidecar, edge_index, edge_attr, num_nodes)
    if g is None:
        g = upload_graph(edge_index, edge_attr, num_nodes)
        save_csr(g, sidecar, edge_index, edge_attr)
    return g


# ---- cache: one CSR per live (edge_index, edge_attr) tensor pair ---------------------------
# Entries hold WEAK references to the tensors they were built from, so a recycled device
# address can never alias a stale CSR; in-place edits are caught through `_version`.
_CACHE: list = []
_CACHE_MAX = 8


def _versions(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], n: int) -> Tuple:
    return (edge_index._version, tuple(edge_index.shape), None if edge_attr is None else edge_attr._version, n)


def get_graph(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], num_nodes: int,
              holder=None) -> GraphCSR:
    """Cached upload.  `holder` (usually the Data object) is checked first, so the steady-state
    cost per forward is one attribute lookup and two identity compares."""
    ver = _versions(edge_index, edge_attr, num_nodes)
    if holder is not None:
        cached = getattr(holder, "_tgcn_graph", None)
        if cached is not None and cached[0]() is edge_index and \
                (cached[1]() if cached[1] is not None else None) is edge_attr and cached[2] == ver:
            return cached[3]
    entry = None
    for e in list(_CACHE):
        ei = e[0]()
        if ei is None:
            _CACHE.remove(e)
            continue
        if ei is edge_index and (e[1]() if e[1] is not None else None) is edge_attr and e[2] == ver:
            entry = e
            break
    if entry is None:
        g = upload_graph(edge_index, edge_attr, num_nodes)
        entry = (weakref.ref(edge_index), weakref.ref(edge_attr) if edge_attr is not None else None, ver, g)
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(0)
        _CACHE.append(entry)
    if holder is not None:
        try:
            object.__setattr__(holder, "_tgcn_graph", entry)
        except Exception:
            pass
    return entry[3]


def clear_cache() -> None:
    _CACHE.clear()
