"""Plan of the hybrid propagation: which blocks of A_hat go to the tensor cores (csrc/spmm_tc.cu) and which
entries stay in the CSR the gather kernel walks (csrc/spmm.cu).  Built once per graph with torch index ops
(device-agnostic: the CPU tests rebuild A_hat from the plan and compare it entry for entry).

The word-word part of a Text2GraphTransformer graph (PMI edges, text2graph.py:156-166) is dominated by hub
words: ranked by degree, the blocks that pair a hub with anything are 5-40 % dense while the matrix as a whole
is < 1 % dense.  Nodes are therefore ranked by degree (rank space is only a view: inputs and outputs stay in
node order -- the operand is permuted while it is packed, partial rows are fetched by rank), the matrix is cut
into 128 x 16 blocks, and a block becomes a dense tile when it holds at least `min_density` * 2048 entries.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

TILE_M, TILE_K = 128, 16
MAX_TILES_PER_UNIT = 96


@dataclass
class TcPlan:
    n_nodes: int
    n_row_blocks: int
    n_col_blocks: int
    rank: torch.Tensor          # int32 [n_rows]: rank of each row (0 = longest)
    perm: torch.Tensor          # int32 [>= n_col_blocks * 16]: column (node) of each column rank, -1 past the end
    tile_rb: torch.Tensor       # int32 [n_tiles] row block of each tile (sorted by (row block, column block))
    tile_kb: torch.Tensor       # int32 [n_tiles]
    A_tiles: torch.Tensor       # fp32 [n_tiles, 128, 16]: the values of A_hat, rows swizzled (split into TF32 hi/lo in the kernel)
    units: torch.Tensor         # int32 [n_units, 4] = {tile_begin, tile_end, slot, row_block}
    slot_ptr: torch.Tensor      # int32 [n_row_blocks + 1]
    n_slots: int
    n_tiles: int
    nnz_dense: int
    remainder: "object"         # GraphCSR of the entries outside the tiles
    min_density: float
    _bufs: Dict[int, tuple] = field(default_factory=dict)

    @property
    def n_units(self) -> int:
        return int(self.units.shape[0])

    def buffers(self, F: int):
        """(Bt workspace, partial-row buffer) for operand width F, allocated once."""
        b = self._bufs.get(F)
        if b is None:
            Fp = (F + 15) // 16 * 16
            dev = self.A_tiles.device
            bt = torch.empty(self.n_col_blocks * 2 * Fp * TILE_K, dtype=torch.float32, device=dev)
            part = torch.empty((max(self.n_slots, 1) * TILE_M, F), dtype=torch.float32, device=dev)
            b = (bt, part)
            self._bufs[F] = b
        return b

    def c_struct(self):
        from . import _native
        p = _native.TcPlanArgs()
        p.A_tiles, p.tile_kb, p.units = self.A_tiles.data_ptr(), self.tile_kb.data_ptr(), self.units.data_ptr()
        p.n_tiles = self.n_tiles
        p.n_units, p.perm, p.n_col_blocks = self.n_units, self.perm.data_ptr(), self.n_col_blocks
        return p


def swizzled_offset(r: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """Float offset of element (row r, column k) inside a [rows][16] fp32 tile stored with the 64-byte swizzle
    (Swizzle<2,4,3>: address bits [4,6) ^= bits [7,9)): the 16-byte chunk k // 4 of the 64-byte row r sits at chunk
    position (k // 4) ^ ((r // 2) % 4)."""
    return r * TILE_K + ((((k >> 2) ^ ((r >> 1) & 3)) << 2) | (k & 3))


def build_tc_plan(graph, min_density: float = 0.05, max_bytes: int = 2 << 30, n_sms: int = 148,
                  width: int = 256, max_operand_bytes: int = 104 << 20) -> Optional[TcPlan]:
    """Returns None when no block of the graph qualifies.

    min_density: a 128 x 16 block becomes a dense tile at >= min_density * 2048 entries.  Measured on B200 (20NG-shape,
        F = 200, profiles/r02_hybrid_spmm.json): 0.05 -> 0.656 ms per propagation, 0.03 -> 0.682, 0.02 -> 0.716 (gather
        kernel alone: 0.879): below ~5 % a tile costs more tensor-core time than the gathers it replaces.
    max_bytes: cap on the tile storage (densest blocks are kept).
    width / max_operand_bytes: the packed operand Bt (2 x 4 bytes x width per rank) must stay L2-resident while the
        tiles stream through, so only column blocks of the first max_operand_bytes / (8 * width) ranks become tiles
        (the hub columns); with Bt in HBM every tile would pull its 2 x 64 x width operand bytes from DRAM."""
    from .graph import GraphCSR
    dev = graph.rowptr.device
    n, n_cols = graph.n_nodes, graph.n_cols            # rectangular for a row shard of the 1D partition (rows = own nodes)
    rows = graph.row_ids()
    cols = graph.colidx.to(torch.int64)
    val = graph.val
    nnz = int(cols.numel())
    # ---- degree ranking: rows by their length, columns by the number of entries they hold ----
    row_len = (graph.rowptr[1:] - graph.rowptr[:-1]).to(torch.int64)
    col_cnt = torch.bincount(cols, minlength=n_cols)
    if n == n_cols:                                     # square matrix: one ranking for both (rows and columns are the same nodes)
        row_order = col_order = torch.sort(row_len + col_cnt, descending=True, stable=True).indices
    else:
        row_order = torch.sort(row_len, descending=True, stable=True).indices
        col_order = torch.sort(col_cnt, descending=True, stable=True).indices
    rank = torch.empty(n, dtype=torch.int64, device=dev)
    rank[row_order] = torch.arange(n, device=dev)
    col_rank = torch.empty(n_cols, dtype=torch.int64, device=dev)
    col_rank[col_order] = torch.arange(n_cols, device=dev)
    n_rb = (n + TILE_M - 1) // TILE_M
    n_kb = (n_cols + TILE_K - 1) // TILE_K
    perm = torch.full(((n_kb * TILE_K + TILE_M - 1) // TILE_M * TILE_M,), -1, dtype=torch.int32, device=dev)
    perm[:n_cols] = col_order.to(torch.int32)
    # ---- block census ----
    rr, rc = rank[rows], col_rank[cols]
    key = (rr >> 7) * n_kb + (rc >> 4)
    # duplicate (row, col) entries (possible in a general COO graph, never emitted by Text2GraphTransformer) cannot share
    # a dense cell: all but the first of each pair stay in the remainder
    full_key = rows * n_cols + cols
    srt = torch.sort(full_key, stable=True)
    dup_sorted = torch.zeros(nnz, dtype=torch.bool, device=dev)
    if nnz > 1:
        dup_sorted[1:] = srt.values[1:] == srt.values[:-1]
    is_dup = torch.zeros(nnz, dtype=torch.bool, device=dev)
    is_dup[srt.indices] = dup_sorted
    Fp = (int(width) + 15) // 16 * 16
    kb_limit = max(1, min(n_kb, int(max_operand_bytes // (2 * Fp * TILE_K * 4))))
    eligible = ~is_dup & ((rc >> 4) < kb_limit)
    ukeys, counts = torch.unique(key[eligible], return_counts=True)
    thr = max(1, int(round(min_density * TILE_M * TILE_K)))
    cand = counts >= thr
    if int(cand.sum().item()) == 0:
        return None
    ck, cc = ukeys[cand], counts[cand]
    max_tiles = max(1, int(max_bytes // (TILE_M * TILE_K * 4)))
    if ck.numel() > max_tiles:                       # memory cap: keep the densest blocks
        top = torch.topk(cc, max_tiles).indices
        ck = torch.sort(ck[top]).values
    sel_keys = ck                                     # sorted by (row block, column block)
    n_tiles = int(sel_keys.numel())
    tile_rb = (sel_keys // n_kb).to(torch.int32)
    tile_kb = (sel_keys % n_kb).to(torch.int32)
    # ---- split the entries ----
    pos = torch.searchsorted(sel_keys, key).clamp_(max=n_tiles - 1)
    dense = (sel_keys[pos] == key) & eligible
    A = torch.zeros((n_tiles, TILE_M * TILE_K), dtype=torch.float32, device=dev)
    t = pos[dense]
    off = swizzled_offset(rr[dense] & (TILE_M - 1), rc[dense] & (TILE_K - 1))
    A.view(-1)[t * (TILE_M * TILE_K) + off] = val[dense]
    keep = ~dense
    r_rows = rows[keep]
    rp = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    rp[1:] = torch.cumsum(torch.bincount(r_rows, minlength=n), 0).to(torch.int32)
    rem = GraphCSR(n, rp, graph.colidx[keep].contiguous(), val[keep].contiguous(), graph.dis, n_cols=n_cols)
    rem._symmetric = False
    # ---- units: <= MAX_TILES_PER_UNIT consecutive tiles of one row block; slots consecutive per row block ----
    per_rb = torch.bincount(tile_rb.to(torch.int64), minlength=n_rb)
    n_u = (per_rb + MAX_TILES_PER_UNIT - 1) // MAX_TILES_PER_UNIT
    slot_ptr = torch.zeros(n_rb + 1, dtype=torch.int64, device=dev)
    slot_ptr[1:] = torch.cumsum(n_u, 0)
    n_slots = int(slot_ptr[-1].item())
    tile_ptr = torch.zeros(n_rb + 1, dtype=torch.int64, device=dev)
    tile_ptr[1:] = torch.cumsum(per_rb, 0)
    u_rb = torch.repeat_interleave(torch.arange(n_rb, device=dev), n_u)
    u_j = torch.arange(n_slots, device=dev) - slot_ptr[u_rb]
    u_begin = tile_ptr[u_rb] + (u_j * per_rb[u_rb]) // n_u[u_rb]
    u_end = tile_ptr[u_rb] + ((u_j + 1) * per_rb[u_rb]) // n_u[u_rb]
    u_slot = torch.arange(n_slots, device=dev)
    # longest units first, dealt to the CTAs in snake order; CTA c walks entries c, c + G, c + 2G, ... of the list
    G = max(1, min(n_sms, n_slots))
    by_len = torch.sort(u_end - u_begin, descending=True, stable=True).indices
    i = torch.arange(n_slots, device=dev)
    wave, p_in = i // G, i % G
    cta = torch.where(wave % 2 == 0, p_in, G - 1 - p_in)
    n_list = int(((n_slots + G - 1) // G) * G)
    units = torch.zeros((n_list, 4), dtype=torch.int32, device=dev)
    units[:, 2] = -1
    dst = wave * G + cta
    units[dst, 0] = u_begin[by_len].to(torch.int32)
    units[dst, 1] = u_end[by_len].to(torch.int32)
    units[dst, 2] = u_slot[by_len].to(torch.int32)
    units[dst, 3] = u_rb[by_len].to(torch.int32)
    n_kb_used = int(tile_kb.max().item()) + 1            # the operand is packed only up to the last column block in use
    return TcPlan(n, n_rb, n_kb_used, rank.to(torch.int32), perm, tile_rb, tile_kb, A.view(n_tiles, TILE_M, TILE_K), units,
                  slot_ptr.to(torch.int32), n_slots, n_tiles, int(dense.sum().item()), rem, float(min_density))
