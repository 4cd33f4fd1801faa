"""CPU emulation of k_spmm_staged's walk over a `StagedPlan` and a host restatement of the chunk
list `tgcn_spmm_plan` builds (csrc/spmm.cu k_plan + the length sort of GraphCSR.plan).  Test
infrastructure only: nothing in the package imports this."""
import torch

from pytextgcn_b200.staged_plan import StagedPlan


def chunk_list(rowptr: torch.Tensor, chunk_nnz: int, row_begin: int = 0, row_end=None, sort: bool = True):
    """(chunks int32 [n,4] = {row, begin, end, slot}, split_rows int32 [m,3] = {row, first_slot, n_slots})
    exactly as tgcn_spmm_plan emits them: rows longer than chunk_nnz split evenly, chunks in row order,
    then (GraphCSR.plan) stably sorted by decreasing length."""
    rp = rowptr.cpu().long()
    row_end = rp.numel() - 1 if row_end is None else row_end
    chunks, split = [], []
    n_slots = 0
    for r in range(row_begin, row_end):
        b, ln = int(rp[r]), int(rp[r + 1] - rp[r])
        nch = max(1, (ln + chunk_nnz - 1) // chunk_nnz)
        if nch == 1:
            chunks.append((r, b, b + ln, -1))
        else:
            per = (ln + nch - 1) // nch
            for c in range(nch):
                cb = b + c * per
                chunks.append((r, cb, min(b + ln, cb + per), n_slots + c))
            split.append((r, n_slots, nch))
            n_slots += nch
    ch = torch.tensor(chunks, dtype=torch.int32).view(-1, 4)
    if sort and ch.shape[0] > 1:
        ch = ch[torch.argsort(ch[:, 2] - ch[:, 1], descending=True, stable=True)].contiguous()
    return ch, torch.tensor(split, dtype=torch.int32).view(-1, 3)


def emulate(plan: StagedPlan, B: torch.Tensor) -> torch.Tensor:
    """CPU walk of the plan in the kernel's order (panel -> consumer warp -> tile -> header -> entries),
    reading operand rows through the staged tile exactly as k_spmm_staged does.  Returns one fp64
    partial row per chunk, [n_chunks, F].  Test infrastructure for the plan layout (small graphs only)."""
    W, RPW, KC = plan.warps_per_panel, plan.rows_per_warp, plan.tile_cols
    stream = plan.stream.cpu()
    ucols = plan.ucols.cpu().long()
    uptr = plan.panel_ucol_ptr.cpu().long()
    wptr = plan.warp_stream_ptr.cpu().long()
    Bc = B.detach().cpu().double()
    out = torch.zeros((plan.n_chunks, Bc.shape[1]), dtype=torch.float64)
    vals = stream[:, 1].contiguous().view(torch.float32).double()
    for p in range(plan.n_panels):
        u0, u1 = int(uptr[p]), int(uptr[p + 1])
        n_tiles = (u1 - u0 + KC - 1) // KC
        for w in range(W):
            q = int(wptr[p * W + w])
            for t in range(n_tiles):
                staged = Bc[ucols[u0 + t * KC: min(u0 + (t + 1) * KC, u1)]]     # the shared-memory stage
                hdr = stream[q]
                q += 1
                for r in range(RPW):
                    cnt = int(hdr[r])
                    vid = (p * W + w) * RPW + r
                    if cnt:
                        assert vid < plan.n_chunks, "entries for a chunk past the end of the list"
                        sl = stream[q:q + cnt, 0].long()
                        assert int(sl.max()) < staged.shape[0], "slot outside the tile"
                        out[vid] += (vals[q:q + cnt, None] * staged[sl]).sum(0)
                    q += cnt
    return out


def rows_from_chunks(partial: torch.Tensor, chunks: torch.Tensor, n_rows: int, row_begin: int = 0) -> torch.Tensor:
    """Adds the per-chunk partial rows of `emulate` into output rows (what finish_row does for split rows)."""
    out = torch.zeros((n_rows, partial.shape[1]), dtype=partial.dtype)
    out.index_add_(0, chunks[:, 0].long() - row_begin, partial)
    return out
