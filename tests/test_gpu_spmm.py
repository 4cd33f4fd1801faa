"""GPU parity: tgcn_spmm (propagation + fused epilogue) against the oracle's edge-wise
gather/scale/scatter-add (GCNConv.propagate restated in oracle/gcn_oracle.py)."""
import pytest
import torch

from helpers import random_graph, rel_err
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5      # north_star: logits within 1e-5 relative in fp32
TOL_BF16 = 1e-2


def _oracle_propagate(ei, w, n, B):
    ei2, w_hat = O.gcn_norm(ei, w, n)
    msg = w_hat.double().view(-1, 1) * B.double().index_select(0, ei2[0])
    return torch.zeros(n, B.shape[1], dtype=torch.float64).index_add_(0, ei2[1], msg)


@pytest.mark.parametrize("F", [4, 8, 20, 32, 64, 100, 128, 200, 256, 220, 512])
def test_plain_spmm_widths(cuda, F):
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    n = 700
    ei, w = random_graph(n, 30000, seed=F)
    g = upload_graph(ei.to(cuda), w.to(cuda), n)
    torch.manual_seed(F)
    B = torch.randn(n, F)
    out, _ = ops.spmm(g, B.to(cuda))
    assert rel_err(out, _oracle_propagate(ei, w, n, B)) < TOL


@pytest.mark.parametrize("chunk", [32, 64, 512])
def test_split_rows_hub(cuda, chunk):
    # hub rows far longer than chunk_nnz go through the partial-row scratch + fix-up kernel
    import numpy as np
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    rng = np.random.default_rng(1)
    n = 2000
    hubs = rng.integers(0, 3, size=9000)
    others = rng.integers(3, n, size=9000)
    key = np.unique(hubs * n + others)
    s, d = key // n, key % n
    ei = torch.from_numpy(np.stack([np.concatenate([s, d]), np.concatenate([d, s])]).astype(np.int64))
    w = torch.from_numpy(np.tile(rng.uniform(0.1, 2.0, size=s.size).astype(np.float32), 2))
    g = upload_graph(ei.to(cuda), w.to(cuda), n)
    plan = g.plan(chunk_nnz=chunk)
    assert plan.n_split_rows >= 3 and plan.max_row_nnz > chunk
    B = torch.randn(n, 200)
    bias = torch.randn(200)
    out, _ = ops.spmm(g, B.to(cuda), plan=plan, bias=bias.to(cuda))
    assert rel_err(out, _oracle_propagate(ei, w, n, B) + bias.double()) < TOL


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("drop", ["none", "mask"])
def test_epilogue_bias_act_dropout_projection(cuda, act, drop):
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    n, F, C = 900, 200, 20
    ei, w = random_graph(n, 40000, seed=7)
    g = upload_graph(ei.to(cuda), w.to(cuda), n)
    torch.manual_seed(3)
    B, bias, W2 = torch.randn(n, F), torch.randn(F), torch.randn(F, C) * 0.2
    keep = torch.rand(n, F) > 0.5
    ref = _oracle_propagate(ei, w, n, B) + bias.double()
    if act:
        ref = torch.relu(ref)
    if drop == "mask":
        ref = ref * keep.double() * 2.0
    out, P = ops.spmm(g, B.to(cuda), bias=bias.to(cuda), act=act,
                      drop_mode=ops.DROP_MASK if drop == "mask" else ops.DROP_NONE, drop_p=0.5,
                      keep_mask=keep.to(torch.uint8).to(cuda), W_proj=W2.to(cuda))
    assert rel_err(out, ref) < TOL
    assert rel_err(P[:, :C], ref @ W2.double()) < TOL


def test_philox_dropout_statistics_and_determinism(cuda):
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    n, F = 1500, 200
    ei, w = random_graph(n, 30000, seed=9)
    g = upload_graph(ei.to(cuda), w.to(cuda), n)
    B = torch.randn(n, F, device=cuda)
    base, _ = ops.spmm(g, B)
    p = 0.7
    a, _ = ops.spmm(g, B, drop_mode=ops.DROP_PHILOX, drop_p=p, philox_seed=123, philox_offset=1)
    b, _ = ops.spmm(g, B, drop_mode=ops.DROP_PHILOX, drop_p=p, philox_seed=123, philox_offset=1)
    c, _ = ops.spmm(g, B, drop_mode=ops.DROP_PHILOX, drop_p=p, philox_seed=123, philox_offset=2)
    assert torch.equal(a, b)                      # same (seed, offset) -> same mask
    assert not torch.equal(a, c)                  # next call -> new mask
    kept = a != 0
    rate = kept.float().mean().item()
    assert abs(rate - (1 - p)) < 0.01             # keep-rate
    assert rel_err(a[kept], base[kept] / (1 - p)) < 1e-6   # survivors scaled by 1/(1-p)
    # per-column / per-row keep rates are unbiased too (no structure from the index mapping)
    assert (kept.float().mean(0) - (1 - p)).abs().max().item() < 0.06
    assert (kept.float().mean(1) - (1 - p)).abs().max().item() < 0.15


@pytest.mark.parametrize("F", [200, 104, 32])
def test_bf16_operand(cuda, F):
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    n = 800
    ei, w = random_graph(n, 30000, seed=21)
    g = upload_graph(ei.to(cuda), w.to(cuda), n)
    B = torch.randn(n, F)
    Bb = ops.cast_bf16(B.to(cuda))
    assert torch.equal(Bb.cpu(), B.to(torch.bfloat16))
    out, _ = ops.spmm(g, Bb, out_dtype=torch.bfloat16)
    ref = _oracle_propagate(ei, w, n, B)
    assert rel_err(out.float(), ref) < TOL_BF16
    # against the same bf16-rounded operand the fp32-accumulated result is tight before the store
    out32, _ = ops.spmm(g, Bb, out_dtype=torch.float32)
    assert rel_err(out32, _oracle_propagate(ei, w, n, B.to(torch.bfloat16).float())) < TOL


def test_row_range_plan(cuda):
    # 1D row partition: a plan over [r0, r1) writes only those rows, at local offsets
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    n, F = 600, 64
    ei, w = random_graph(n, 20000, seed=4)
    g = upload_graph(ei.to(cuda), w.to(cuda), n)
    B = torch.randn(n, F)
    ref = _oracle_propagate(ei, w, n, B)
    parts = []
    for r0, r1 in [(0, 150), (150, 151), (151, 600)]:
        out, _ = ops.spmm(g, B.to(cuda), plan=g.plan(r0, r1, chunk_nnz=64))
        assert out.shape[0] == r1 - r0
        parts.append(out.cpu())
    assert rel_err(torch.cat(parts), ref) < TOL


def test_bad_arguments_raise(cuda):
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    ei, w = random_graph(50, 300, seed=0)
    g = upload_graph(ei.to(cuda), w.to(cuda), 50)
    with pytest.raises(RuntimeError):
        ops.spmm(g, torch.randn(50, 6, device=cuda))            # F not a multiple of 4
    with pytest.raises(RuntimeError):
        ops.spmm(g, torch.randn(40, 8, device=cuda))            # too few rows
    with pytest.raises(RuntimeError):
        ops.spmm(g, torch.randn(50, 8))                         # CPU tensor
