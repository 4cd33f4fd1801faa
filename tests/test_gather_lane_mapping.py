"""Lane arithmetic of the gather kernel k_spmm (csrc/spmm.cu) restated in numpy: which lane handles which
non-zero / which piece of the gathered row, for the shuffle-fed loop of the shipped build and for the build
variants (TGCN_SPMM_EXACTLPR: lanes per row not a power of two; TGCN_SPMM_CVPACK: one 8-byte pair load per
non-zero).  Every non-zero of a chunk must be counted exactly once for every chunk length and row width."""
import numpy as np
import pytest


def _walk(LPR, n_entries, U, cvpack):
    NZP = 32 // LPR
    rng = np.random.default_rng(LPR * 100 + n_entries)
    cols, vals = rng.integers(0, 50, n_entries), rng.standard_normal(n_entries)
    B = rng.standard_normal((50, LPR))                      # one value per lane of a row
    acc = np.zeros(32)
    begin, end = 7, 7 + n_entries
    gcol, gval = np.zeros(end + 64, int), np.zeros(end + 64)
    gcol[begin:end], gval[begin:end] = cols, vals
    lanes = np.arange(32)
    sub, l = lanes // LPR, lanes % LPR
    if not cvpack:
        for base in range(begin, end, 32):                  # 32 entries per coalesced load, broadcast by shuffle
            k = base + lanes
            mc, mv = np.where(k < end, gcol[k], 0), np.where(k < end, gval[k], 0.0)
            cnt = min(32, end - base)
            for j in range(0, cnt, NZP * U):
                for u in range(U):
                    srcl = j + u * NZP + sub
                    v = mv[srcl & 31].copy()
                    v[srcl >= cnt] = 0
                    if 32 % LPR:
                        v[sub >= NZP] = 0                   # lanes past the last whole group idle
                    acc += v * B[mc[srcl & 31], l]
    else:
        for base in range(begin, end, NZP * U):             # one pair load per non-zero and lane group
            for u in range(U):
                k = base + u * NZP + sub
                ok = (k < end) & ((32 % LPR == 0) | (sub < NZP))
                acc += np.where(ok, gval[k], 0.0) * B[np.where(ok, gcol[k], 0), l]
    if LPR & (LPR - 1) == 0:                                # shuffle-down tree
        o = 16
        while o >= LPR:
            acc = acc + np.concatenate([acc[o:], acc[32 - o:]])
            o //= 2
        out = acc[:LPR]
    else:                                                   # group g of lane l sits at lane l + g*LPR
        out = np.array([acc[x] + sum(acc[(x + g * LPR) & 31] for g in range(1, NZP)) for x in range(LPR)])
    return out, (vals[:, None] * B[cols]).sum(0)


@pytest.mark.parametrize("cvpack", [False, True])
@pytest.mark.parametrize("LPR,U", [(2, 4), (3, 4), (4, 4), (5, 4), (6, 4), (7, 4), (8, 4), (16, 4), (32, 8), (32, 2)])
def test_every_nonzero_counted_once(LPR, U, cvpack):
    for n in [0, 1, 5, 23, 24, 25, 31, 32, 33, 64, 100, 257]:
        out, ref = _walk(LPR, n, U, cvpack)
        assert np.abs(out - ref).max() < 1e-9, (LPR, n)
