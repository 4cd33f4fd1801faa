"""Layout of the staged-SpMM plan (pytextgcn_b200/staged_plan.py), checked on the CPU by walking it
exactly as k_spmm_staged does (tests/staged_emulator.py) and comparing with the oracle's
A_hat @ B (oracle/gcn_oracle.py, restating [PyG-1.6.3] gcn_norm + propagate; models.py:20)."""
import numpy as np
import pytest
import torch

from helpers import random_graph, rel_err
from oracle import gcn_oracle as O
from pytextgcn_b200.staged_plan import build_staged_plan, STREAM_PAD
from pytextgcn_b200.synthetic import make_graph
from staged_emulator import chunk_list, emulate, rows_from_chunks


def _csr(ei, w, n):
    rowptr, colidx, val = O.csr_from_gcn_norm(ei, w, n)[:3]
    return rowptr.to(torch.int32), colidx.to(torch.int32), val


def _dense_ref(rowptr, colidx, val, B):
    n = rowptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n), (rowptr[1:] - rowptr[:-1]).long())
    out = torch.zeros((n, B.shape[1]), dtype=torch.float64)
    out.index_add_(0, rows, val.double()[:, None] * B.double()[colidx.long()])
    return out


@pytest.mark.parametrize("W,RPW,KC,chunk_nnz", [(28, 1, 64, 64), (28, 2, 64, 32), (5, 1, 7, 32), (3, 2, 128, 2048), (31, 1, 1, 40)])
def test_plan_walk_reproduces_spmm(W, RPW, KC, chunk_nnz):
    n = 300
    ei, w = random_graph(n, 6000, seed=3, duplicates=40, self_loops=5, isolated=4)
    rowptr, colidx, val = _csr(ei, w, n)
    chunks, split = chunk_list(rowptr, chunk_nnz)
    plan = build_staged_plan(colidx, val, chunks, n, warps_per_panel=W, rows_per_warp=RPW, tile_cols=KC)
    B = torch.randn(n, 24, generator=torch.Generator().manual_seed(1))
    got = rows_from_chunks(emulate(plan, B), chunks, n)
    assert rel_err(got, _dense_ref(rowptr, colidx, val, B)) < 1e-12
    # structural invariants the kernel relies on
    R = W * RPW
    assert plan.n_panels == (chunks.shape[0] + R - 1) // R
    assert plan.stream.shape[0] == plan.stream_len + STREAM_PAD
    assert plan.warp_stream_ptr.numel() == plan.n_panels * W
    up = plan.panel_ucol_ptr.long()
    for p in range(plan.n_panels):
        u = plan.ucols[up[p]:up[p + 1]].long()
        assert bool((u[1:] > u[:-1]).all()), "union columns of a panel must be strictly ascending"
    # every non-zero appears exactly once; every tile of every warp has a header
    n_tiles = (up[1:] - up[:-1] + KC - 1) // KC
    assert plan.stream_len == int(n_tiles.sum()) * W + int(rowptr[-1])
    assert plan.gathered_rows() <= plan.nnz


def test_plan_on_a_row_shard_and_textgcn_graph():
    g = make_graph("tiny", seed=2)
    n = g.x.shape[0]
    rowptr, colidx, val = _csr(g.edge_index, g.edge_attr, n)
    lo, hi = 37, 171                                    # a row range, as the 1D row partition uses
    chunks, _ = chunk_list(rowptr, 32, lo, hi)
    plan = build_staged_plan(colidx, val, chunks, n, warps_per_panel=6, rows_per_warp=2, tile_cols=16)
    B = torch.randn(n, 8, generator=torch.Generator().manual_seed(5))
    got = rows_from_chunks(emulate(plan, B), chunks, hi - lo, row_begin=lo)
    assert rel_err(got, _dense_ref(rowptr, colidx, val, B)[lo:hi]) < 1e-12
    # popular word columns are shared inside a panel: fewer staged rows than non-zeros
    assert plan.gathered_rows() < plan.nnz


def test_empty_and_degenerate_plans():
    z = torch.zeros(0, dtype=torch.int32)
    p = build_staged_plan(z, torch.zeros(0), torch.zeros((0, 4), dtype=torch.int32), 10)
    assert p.n_panels == 0 and p.stream.shape[0] == STREAM_PAD
    # rows without any entry (a CSR row range of a graph without self loops): headers only
    rowptr = torch.tensor([0, 0, 2, 2, 5], dtype=torch.int32)
    colidx = torch.tensor([1, 3, 0, 1, 2], dtype=torch.int32)
    val = torch.tensor([1., 2., 3., 4., 5.])
    chunks, _ = chunk_list(rowptr, 32)
    plan = build_staged_plan(colidx, val, chunks, 4, warps_per_panel=2, rows_per_warp=1, tile_cols=2)
    B = torch.arange(8, dtype=torch.float32).view(4, 2)
    got = rows_from_chunks(emulate(plan, B), chunks, 4)
    assert torch.equal(got, _dense_ref(rowptr, colidx, val, B))
    with pytest.raises(ValueError):
        build_staged_plan(colidx, val, chunks, 4, rows_per_warp=3)
    with pytest.raises(ValueError):
        build_staged_plan(colidx, val, chunks, 4, tile_cols=129)


def test_plan_property_random_csr():
    """Random CSRs (empty rows, duplicate columns, hub rows) x random plan shapes: the kernel's walk over the
    plan always reproduces the row sums."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 60), st.integers(0, 10 ** 6), st.integers(1, 31), st.sampled_from([1, 2]),
           st.integers(1, 128), st.sampled_from([32, 33, 64, 500]))
    def check(n_rows, n_cols, seed, W, RPW, KC, chunk_nnz):
        rng = np.random.default_rng(seed)
        lens = rng.integers(0, 6, size=n_rows)
        lens[rng.integers(0, n_rows)] = rng.integers(0, 300)         # one hub row, possibly split
        rowptr = torch.zeros(n_rows + 1, dtype=torch.int32)
        rowptr[1:] = torch.from_numpy(np.cumsum(lens)).to(torch.int32)
        nnz = int(rowptr[-1])
        colidx = torch.from_numpy(rng.integers(0, n_cols, size=nnz)).to(torch.int32)     # duplicates allowed
        val = torch.from_numpy(rng.standard_normal(nnz).astype(np.float32))
        chunks, _ = chunk_list(rowptr, chunk_nnz)
        plan = build_staged_plan(colidx, val, chunks, n_cols, warps_per_panel=W, rows_per_warp=RPW, tile_cols=KC)
        B = torch.from_numpy(rng.standard_normal((n_cols, 4)).astype(np.float32))
        got = rows_from_chunks(emulate(plan, B), chunks, n_rows)
        ref = _dense_ref(rowptr, colidx, val, B)
        assert float((got - ref).abs().max()) <= 1e-9 * max(1.0, float(ref.abs().max()))
    check()
