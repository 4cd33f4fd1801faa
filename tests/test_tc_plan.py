"""Host logic of the hybrid propagation (pytextgcn_b200/tc_plan.py) on the CPU: the dense tiles (un-swizzled)
plus the remainder CSR must rebuild A_hat entry for entry, and every tile must belong to exactly one unit."""
import pytest
import torch

from oracle import gcn_oracle as O
from pytextgcn_b200.graph import GraphCSR
from pytextgcn_b200.synthetic import make_graph
from pytextgcn_b200.tc_plan import TILE_K, TILE_M, build_tc_plan, swizzled_offset


def _csr(name, seed):
    g = make_graph(name, seed=seed)
    n = int(g.x.shape[0])
    rowptr, col, val, dis, _ = O.csr_from_gcn_norm(g.edge_index, g.edge_attr, n)
    return n, rowptr, col, val, GraphCSR(n, rowptr.to(torch.int32), col.to(torch.int32), val, dis)


def _dense(n, rowptr, col, val):
    rows = torch.repeat_interleave(torch.arange(n), (rowptr[1:] - rowptr[:-1]).long())
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((rows, col.long()), val.double(), accumulate=True)
    return A


@pytest.mark.parametrize("name,density,n_sms", [("tiny", 0.002, 4), ("small", 0.01, 148), ("small", 0.2, 7)])
def test_plan_rebuilds_a_hat_exactly(name, density, n_sms):
    n, rowptr, col, val, gr = _csr(name, 0)
    tc = build_tc_plan(gr, min_density=density, n_sms=n_sms)
    assert tc is not None and tc.nnz_dense + tc.remainder.nnz == gr.nnz
    perm = tc.perm.long()
    assert torch.equal(tc.rank.long()[perm[:n]], torch.arange(n)) and bool((perm[n:] == -1).all())      # square: one ranking
    off = swizzled_offset(torch.arange(TILE_M).view(-1, 1), torch.arange(TILE_K).view(1, -1))
    assert sorted(off.view(-1).tolist()) == list(range(TILE_M * TILE_K))          # the swizzle is a permutation of the tile
    A = _dense(n, tc.remainder.rowptr, tc.remainder.colidx, tc.remainder.val)
    for t in range(tc.n_tiles):
        tile = tc.A_tiles[t].reshape(-1)[off].double()
        rn = perm[int(tc.tile_rb[t]) * TILE_M:][:TILE_M]
        cn = perm[int(tc.tile_kb[t]) * TILE_K:][:TILE_K]
        vr, vc = rn >= 0, cn >= 0
        assert float(tile[~vr].abs().sum()) == 0 and float(tile[:, ~vc].abs().sum()) == 0
        A[rn[vr].view(-1, 1), cn[vc].view(1, -1)] += tile[vr][:, vc]
    assert torch.equal(A, _dense(n, rowptr, col, val))
    # units: every tile in exactly one unit, of its own row block, with a slot of that row block; list dealt to n_sms CTAs
    cover = torch.zeros(tc.n_tiles, dtype=torch.int32)
    for b, e, s, rb in tc.units.tolist():
        if e > b:
            cover[b:e] += 1
            assert bool((tc.tile_rb[b:e] == rb).all()) and int(tc.slot_ptr[rb]) <= s < int(tc.slot_ptr[rb + 1]) and e - b <= 96
    assert bool((cover == 1).all())
    assert sorted(set(tc.units[:, 2].tolist()) - {-1}) == list(range(tc.n_slots))
    assert tc.n_units % min(n_sms, tc.n_slots) == 0


def test_duplicate_entries_stay_in_the_remainder():
    n = 200
    rowptr = torch.zeros(n + 1, dtype=torch.int32)
    rowptr[1:] = 3
    rowptr = torch.cumsum(rowptr, 0).to(torch.int32)
    col = torch.tensor([0, 1, 1] * n, dtype=torch.int32)                     # (r, 1) twice in every row
    val = torch.rand(3 * n)
    gr = GraphCSR(n, rowptr, col, val, None)
    tc = build_tc_plan(gr, min_density=1 / 4096, n_sms=4)
    assert tc is not None and tc.nnz_dense == 2 * n and tc.remainder.nnz == n


def test_no_dense_block_gives_no_plan():
    n, rowptr, col, val, gr = _csr("tiny", 1)
    assert build_tc_plan(gr, min_density=0.9) is None


def test_operand_cap_restricts_tiles_to_hub_column_blocks():
    n, rowptr, col, val, gr = _csr("small", 2)
    full = build_tc_plan(gr, min_density=0.01, n_sms=8, width=64)
    capped = build_tc_plan(gr, min_density=0.01, n_sms=8, width=64, max_operand_bytes=10 * 2 * 64 * TILE_K * 4)   # 10 column blocks
    assert capped is not None and capped.n_col_blocks <= 10 and int(capped.tile_kb.max()) < 10
    assert 0 < capped.nnz_dense < full.nnz_dense and capped.nnz_dense + capped.remainder.nnz == gr.nnz


def test_plan_of_a_row_shard_rebuilds_the_shard():
    """Rectangular matrix (rows of one rank of the 1D partition, columns in the padded id space): separate row and
    column rankings; tiles + remainder must rebuild the shard exactly."""
    from pytextgcn_b200.dist import RowPartition, shard_csr
    n, rowptr, col, val, gr = _csr("small", 3)
    part = RowPartition(rowptr[1:] - rowptr[:-1], 3)
    rp, ci, v = shard_csr(rowptr, col, val, part, 1)
    shard = GraphCSR(part.n_loc, rp, ci, v, None, n_cols=part.n_pad)
    tc = build_tc_plan(shard, min_density=0.02, n_sms=5, width=64)
    assert tc is not None and tc.nnz_dense + tc.remainder.nnz == shard.nnz and tc.remainder.n_cols == part.n_pad
    perm, rrank = tc.perm.long(), tc.rank.long()
    row_of_rank = torch.empty(part.n_loc, dtype=torch.int64)
    row_of_rank[rrank] = torch.arange(part.n_loc)
    off = swizzled_offset(torch.arange(TILE_M).view(-1, 1), torch.arange(TILE_K).view(1, -1))
    rows_r = torch.repeat_interleave(torch.arange(part.n_loc), (tc.remainder.rowptr[1:] - tc.remainder.rowptr[:-1]).long())
    A = torch.zeros(part.n_loc, part.n_pad, dtype=torch.float64)
    A.index_put_((rows_r, tc.remainder.colidx.long()), tc.remainder.val.double(), accumulate=True)
    for t in range(tc.n_tiles):
        tile = tc.A_tiles[t].reshape(-1)[off].double()
        r0, k0 = int(tc.tile_rb[t]) * TILE_M, int(tc.tile_kb[t]) * TILE_K
        nr = min(TILE_M, part.n_loc - r0)
        cn = perm[k0:k0 + TILE_K]
        vc = cn >= 0
        assert float(tile[nr:].abs().sum()) == 0 and float(tile[:, ~vc].abs().sum()) == 0
        A[row_of_rank[r0:r0 + nr].view(-1, 1), cn[vc].view(1, -1)] += tile[:nr][:, vc]
    ref = torch.zeros(part.n_loc, part.n_pad, dtype=torch.float64)
    rows_s = torch.repeat_interleave(torch.arange(part.n_loc), (rp[1:] - rp[:-1]).long())
    ref.index_put_((rows_s, ci.long()), v.double(), accumulate=True)
    assert torch.equal(A, ref)
