"""Plan for the shared-memory staged panel SpMM (csrc/spmm_staged.cu, tgcn_spmm_staged).

Same operation as `tgcn_spmm` (GCNConv.propagate, textgcn/lib/models.py:20), different data
movement: one CTA owns a PANEL of consecutive chunks of the length-sorted chunk list and walks
the sorted union of the columns those chunks touch, tile by tile, with the operand rows of a
tile staged in shared memory.  This module turns (CSR, chunk list) into the arrays that kernel
reads -- layout documented in include/textgcn_b200.h next to `tgcn_staged_plan`.

The builder is written with torch index ops only, so it runs on the device the CSR lives on
(one-off, at graph upload) and, unchanged, on CPU tensors -- which is how the layout is tested
without a GPU: tests/staged_emulator.py walks the plan exactly as the kernel does.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _native

STREAM_PAD = 64          # pairs appended to the entry stream (the kernel prefetches past a warp's last entry)
MAX_TILE_COLS = 128      # STAGED_MAX_TILE_COLS in spmm_staged.cu
PROD_BULK, PROD_LDGSTS = 0, 1


@dataclass
class StagedPlan:
    """Device arrays of `tgcn_staged_plan` for one (chunk list, warps_per_panel, rows_per_warp, tile_cols)."""
    n_chunks: int
    n_panels: int
    warps_per_panel: int
    rows_per_warp: int
    tile_cols: int
    panel_ucol_ptr: torch.Tensor     # int32 [n_panels + 1]
    ucols: torch.Tensor              # int32 [n_ucols]
    warp_stream_ptr: torch.Tensor    # int64 [n_panels * warps_per_panel]
    stream: torch.Tensor             # int32 [stream_len + STREAM_PAD, 2]
    stream_len: int
    nnz: int

    @property
    def rows_per_panel(self) -> int:
        return self.warps_per_panel * self.rows_per_warp

    def gathered_rows(self) -> int:
        """Operand rows copied L2 -> shared memory per launch (the unstaged kernel moves `nnz` of them)."""
        return int(self.ucols.numel())

    def bytes(self) -> int:
        return int(self.stream.numel() * 4 + self.ucols.numel() * 4 + self.panel_ucol_ptr.numel() * 4
                   + self.warp_stream_ptr.numel() * 8)


def build_staged_plan(colidx: torch.Tensor, val: torch.Tensor, chunks: torch.Tensor, n_cols: int, *,
                      warps_per_panel: int = 28, rows_per_warp: int = 2, tile_cols: int = 64) -> StagedPlan:
    """colidx int32 [nnz], val fp32 [nnz]: the CSR arrays the chunk list indexes; chunks int32
    [n_chunks, 4] = {row, begin, end, slot} in the order the kernel will use (tgcn_spmm_plan, sorted)."""
    if rows_per_warp not in (1, 2):
        raise ValueError("rows_per_warp must be 1 or 2")
    if not (1 <= tile_cols <= MAX_TILE_COLS):
        raise ValueError(f"tile_cols must be in [1, {MAX_TILE_COLS}]")
    if not (1 <= warps_per_panel <= 31):
        raise ValueError("warps_per_panel must be in [1, 31]")
    dev = colidx.device
    i64 = torch.int64
    W, RPW, KC = int(warps_per_panel), int(rows_per_warp), int(tile_cols)
    R = W * RPW
    n_chunks = int(chunks.shape[0])
    n_panels = (n_chunks + R - 1) // R
    if n_chunks == 0:
        z32 = torch.zeros(1, dtype=torch.int32, device=dev)
        return StagedPlan(0, 0, W, RPW, KC, z32, torch.zeros(0, dtype=torch.int32, device=dev),
                          torch.zeros(0, dtype=i64, device=dev),
                          torch.zeros((STREAM_PAD, 2), dtype=torch.int32, device=dev), 0, 0)

    begin = chunks[:, 1].to(i64)
    lens = chunks[:, 2].to(i64) - begin
    total = int(lens.sum().item())
    # entry k of the concatenated chunk ranges -> its chunk (position in the list) and its CSR slot
    vid = torch.repeat_interleave(torch.arange(n_chunks, device=dev, dtype=i64), lens)
    first = torch.cumsum(lens, 0) - lens
    idx = torch.arange(total, device=dev, dtype=i64) - first[vid] + begin[vid]
    col = colidx[idx].to(i64)
    v = val[idx]

    # union of columns per panel, ascending: unique over (panel, column)
    panel = vid // R
    ukeys, inv = torch.unique(panel * n_cols + col, sorted=True, return_inverse=True)
    ucols = (ukeys % n_cols).to(torch.int32)
    bounds = torch.arange(n_panels + 1, device=dev, dtype=i64) * n_cols
    panel_ucol_ptr = torch.searchsorted(ukeys, bounds).to(i64)
    if int(panel_ucol_ptr[-1].item()) >= 2 ** 31:
        raise RuntimeError("staged plan: more than 2^31 staged rows")
    pos = inv - panel_ucol_ptr[panel]                  # position of the entry's column in its panel's union
    tile = pos // KC
    slot = pos - tile * KC

    n_tiles = (panel_ucol_ptr[1:] - panel_ucol_ptr[:-1] + KC - 1) // KC          # per panel
    warp_tiles = torch.repeat_interleave(n_tiles, W)                               # per consumer warp
    hdr_base = torch.cumsum(warp_tiles, 0) - warp_tiles                            # first header of each warp
    n_hdr = int(warp_tiles.sum().item())

    # stream order: (warp, tile, chunk of the warp, original order)
    warp = vid // RPW
    r = vid - warp * RPW
    hdr_of_entry = hdr_base[warp] + tile
    order = torch.sort(hdr_of_entry * RPW + r, stable=True).indices
    hdr_sorted = hdr_of_entry[order]
    counts = torch.bincount(hdr_of_entry * RPW + r, minlength=n_hdr * RPW).view(n_hdr, RPW)
    per_hdr = counts.sum(1)
    hdr_pos = torch.arange(n_hdr, device=dev, dtype=i64) + torch.cumsum(per_hdr, 0) - per_hdr
    stream_len = n_hdr + total
    stream = torch.zeros((stream_len + STREAM_PAD, 2), dtype=torch.int32, device=dev)
    stream[hdr_pos, 0] = counts[:, 0].to(torch.int32)
    if RPW == 2:
        stream[hdr_pos, 1] = counts[:, 1].to(torch.int32)
    ent_pos = torch.arange(total, device=dev, dtype=i64) + hdr_sorted + 1
    stream[ent_pos, 0] = slot[order].to(torch.int32)
    stream[ent_pos, 1] = v[order].contiguous().view(torch.int32)
    # start of every warp's stream (a warp of a panel without tiles has an empty stream)
    hdr_pos_ext = torch.cat([hdr_pos, torch.tensor([stream_len], device=dev, dtype=i64)])
    warp_stream_ptr = hdr_pos_ext[hdr_base].contiguous()
    return StagedPlan(n_chunks, n_panels, W, RPW, KC, panel_ucol_ptr.to(torch.int32).contiguous(), ucols.contiguous(),
                      warp_stream_ptr, stream, stream_len, total)


def c_plan(plan: StagedPlan, n_producers: int = 4, producer_mode: int = PROD_BULK) -> _native.StagedPlanArgs:
    """ctypes image of `tgcn_staged_plan` (include/textgcn_b200.h)."""
    if plan.warps_per_panel + n_producers > 32:
        raise ValueError("warps_per_panel + n_producers must be at most 32 (one CTA of at most 1024 threads)")
    c = _native.StagedPlanArgs()
    c.panel_ucol_ptr, c.ucols = plan.panel_ucol_ptr.data_ptr(), plan.ucols.data_ptr()
    c.warp_stream_ptr, c.stream = plan.warp_stream_ptr.data_ptr(), plan.stream.data_ptr()
    c.n_panels, c.warps_per_panel = plan.n_panels, plan.warps_per_panel
    c.rows_per_warp, c.tile_cols = plan.rows_per_warp, plan.tile_cols
    c.n_producers, c.producer_mode = int(n_producers), int(producer_mode)
    return c
