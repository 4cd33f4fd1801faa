"""CPU-side report for the staged panel SpMM: how many operand rows a launch copies L2 -> shared memory
for a named synthetic shape, per plan configuration, against the nnz rows the unstaged kernel gathers.
Runs without a GPU (the plan builder is device-agnostic torch code; the chunk list is restated on the host).

    python tools/staged_reuse.py [shape=20ng] [chunk_nnz=auto]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import gcn_oracle as O  # noqa: E402
from pytextgcn_b200.graph import auto_chunk_nnz  # noqa: E402
from pytextgcn_b200.staged_plan import build_staged_plan  # noqa: E402
from pytextgcn_b200.synthetic import SHAPES, make_graph  # noqa: E402


def chunk_list_vectorised(rowptr: torch.Tensor, chunk_nnz: int) -> torch.Tensor:
    """tgcn_spmm_plan + the length sort of GraphCSR.plan, vectorised (slots are not needed here)."""
    lens = (rowptr[1:] - rowptr[:-1]).long()
    nch = torch.clamp((lens + chunk_nnz - 1) // chunk_nnz, min=1)
    row = torch.repeat_interleave(torch.arange(lens.numel()), nch)
    c = torch.arange(row.numel()) - (torch.cumsum(nch, 0) - nch)[row]
    per = (lens[row] + nch[row] - 1) // nch[row]
    b = rowptr[:-1].long()[row] + c * per
    e = torch.minimum(rowptr[1:].long()[row], b + per)
    ch = torch.stack([row, b, e, torch.full_like(row, -1)], 1).to(torch.int32)
    return ch[torch.argsort(ch[:, 2] - ch[:, 1], descending=True, stable=True)].contiguous()


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "20ng"
    g = make_graph(shape)
    n = g.x.shape[0]
    rowptr, colidx, val = O.csr_from_gcn_norm(g.edge_index, g.edge_attr, n)[:3]
    nnz = int(rowptr[-1])
    chunk_nnz = int(sys.argv[2]) if len(sys.argv) > 2 else auto_chunk_nnz(nnz)
    chunks = chunk_list_vectorised(rowptr, chunk_nnz)
    F = SHAPES[shape].hidden
    print(f"{shape}: {n} nodes, nnz {nnz}, chunk_nnz {chunk_nnz}, {chunks.shape[0]} chunks, F {F}")
    for W, RPW, KC in [(28, 1, 64), (28, 2, 64), (30, 1, 64), (16, 1, 64), (28, 1, 32), (28, 1, 128)]:
        t = time.time()
        p = build_staged_plan(colidx.to(torch.int32), val, chunks, n, warps_per_panel=W, rows_per_warp=RPW, tile_cols=KC)
        up = p.panel_ucol_ptr.long()
        tiles = int(((up[1:] - up[:-1] + KC - 1) // KC).sum())
        print(f"  W={W} RPW={RPW} KC={KC}: panels {p.n_panels}, staged rows {p.gathered_rows()} = "
              f"{p.gathered_rows() / nnz:.3f} x nnz ({p.gathered_rows() * F * 4 / 1e9:.2f} GB vs {nnz * F * 4 / 1e9:.2f} GB), "
              f"tiles {tiles}, entries per (tile, warp) {nnz / max(tiles * W, 1):.1f}, plan {p.bytes() / 1e6:.0f} MB, "
              f"built in {time.time() - t:.1f} s")


if __name__ == "__main__":
    main()
