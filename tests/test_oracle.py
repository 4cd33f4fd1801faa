"""CPU checks of the oracle itself (oracle/gcn_oracle.py): the edge-wise PyG-1.6.3 restatement
against an independent dense fp64 formulation of A_hat = D^-1/2 (A+I) D^-1/2, the CSR derived
from it, and autograd gradcheck of the whole forward."""
import numpy as np
import pytest
import torch

from helpers import karate_graph, random_graph, rel_err
from oracle import gcn_oracle as O


def _dense_from_csr(rowptr, col, val, n):
    A = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for k in range(int(rowptr[i]), int(rowptr[i + 1])):
            A[i, int(col[k])] += float(val[k])
    return A


@pytest.mark.parametrize("case", ["sym", "directed", "loops", "dups", "isolated"])
def test_gcn_norm_matches_dense_formulation(case):
    n = 40
    kw = dict(sym=dict(), directed=dict(symmetric=False), loops=dict(self_loops=5),
              dups=dict(duplicates=20), isolated=dict(isolated=6))[case]
    ei, w = random_graph(n, 300, seed=3, **kw)
    ei2, w_hat = O.gcn_norm(ei, w, n)
    assert ei2.shape[1] == w_hat.numel()
    A = torch.zeros(n, n, dtype=torch.float64)
    A.index_put_((ei2[1], ei2[0]), w_hat.double(), accumulate=True)
    ref = O.dense_ahat_fp64(ei, w, n)
    assert rel_err(A, ref) < 1e-6


def test_self_loops_appended_last_with_weight_one():
    ei, w = random_graph(10, 30, seed=1)
    ei2, w2 = O.add_remaining_self_loops(ei, w, 1.0, 10)
    assert torch.equal(ei2[:, -10:], torch.arange(10).repeat(2, 1))
    assert torch.all(w2[-10:] == 1.0)
    assert torch.equal(ei2[:, :-10], ei)


def test_existing_self_loop_weight_is_kept():
    ei = torch.tensor([[0, 1, 2, 1], [1, 0, 2, 2]])
    w = torch.tensor([0.5, 0.5, 3.0, 0.25])
    ei2, w2 = O.add_remaining_self_loops(ei, w, 1.0, 3)
    assert ei2.shape[1] == 3 + 3
    assert w2[-3:].tolist() == [1.0, 1.0, 3.0]


def test_degree_is_sequential_fp32_sum_and_ieee_rsqrt():
    # the two torch-CPU behaviours the device build must reproduce bit for bit
    n = 50
    ei, w = random_graph(n, 4000, seed=5, weight_range=(1e-3, 7.0))
    ei2, w2 = O.add_remaining_self_loops(ei, w, 1.0, n)
    deg = torch.zeros(n).scatter_add_(0, ei2[1], w2)
    seq = np.zeros(n, dtype=np.float32)
    for c, v in zip(ei2[1].tolist(), w2.numpy()):
        seq[c] = np.float32(seq[c] + v)
    assert np.array_equal(deg.numpy(), seq)
    dis = deg.clone().pow_(-0.5)
    ieee = (np.float32(1.0) / np.sqrt(seq)).astype(np.float32)
    assert np.array_equal(dis.numpy(), ieee)


def test_csr_from_gcn_norm_layout():
    n = 30
    ei, w = random_graph(n, 200, seed=2, self_loops=3, duplicates=10)
    rowptr, col, val, dis, perm = O.csr_from_gcn_norm(ei, w, n)
    ei2, w_hat = O.gcn_norm(ei, w, n)
    assert rowptr[-1] == w_hat.numel()
    for i in range(n):
        seg = perm[rowptr[i]:rowptr[i + 1]]
        assert torch.all(ei2[1][seg] == i)
        assert torch.all(seg[1:] > seg[:-1])                 # original edge order inside the row
        assert int(col[rowptr[i + 1] - 1]) == i              # self loop last
    assert rel_err(_dense_from_csr(rowptr, col, val, n), O.dense_ahat_fp64(ei, w, n)) < 1e-6


@pytest.mark.parametrize("relu", [False, True])
def test_forward_matches_dense_fp64(relu):
    g = karate_graph()
    n = 34
    torch.manual_seed(0)
    W = [torch.randn(n, 16) * 0.3, torch.randn(16, 4) * 0.3]
    b = [torch.randn(16) * 0.1, torch.randn(4) * 0.1]
    mask = [torch.rand(n, 16) > 0.5]
    out = O.gcn_forward(g.x, g.edge_index, g.edge_attr, W, b, p=0.5, training=True, drop_masks=mask, relu=relu)
    ref = O.dense_forward_fp64(torch.eye(n), g.edge_index, g.edge_attr, W, b, p=0.5, drop_masks=mask, relu=relu)
    assert rel_err(out, ref) < 1e-5


def test_hierarchy_features_forward():
    n_vocab, n_docs, cprev = 12, 9, 3
    n = n_vocab + n_docs
    ei, w = random_graph(n, 120, seed=4)
    hf = torch.nn.functional.one_hot(torch.arange(n_docs) % cprev, cprev).float()
    x = O.sparse_identity_features(n, hf, n_vocab)
    assert tuple(x.shape) == (n, n + cprev)
    xd = x.to_dense()
    assert torch.equal(xd[:, :n], torch.eye(n))
    assert torch.equal(xd[n_vocab:, n:], hf) and torch.all(xd[:n_vocab, n:] == 0)
    W = [torch.randn(n + cprev, 8), torch.randn(8, 3)]
    b = [torch.zeros(8), torch.zeros(3)]
    out = O.gcn_forward(x, ei, w, W, b)
    ref = O.dense_forward_fp64(xd, ei, w, W, b)
    assert rel_err(out, ref) < 1e-5


def test_gradcheck_fp64():
    g = karate_graph()
    n = 34
    torch.manual_seed(1)
    W1 = (torch.randn(n, 6, dtype=torch.float64) * 0.3).requires_grad_()
    b1 = (torch.randn(6, dtype=torch.float64) * 0.1).requires_grad_()
    W2 = (torch.randn(6, 3, dtype=torch.float64) * 0.3).requires_grad_()
    b2 = (torch.randn(3, dtype=torch.float64) * 0.1).requires_grad_()
    x = g.x.to(torch.float64)
    ew = g.edge_attr.double()
    y = g.y % 3

    def f(W1, b1, W2, b2):
        z = O.gcn_forward(x, g.edge_index, ew, [W1, W2], [b1, b2])
        return O.masked_cross_entropy(z, y, g.train_mask)

    assert torch.autograd.gradcheck(f, (W1, b1, W2, b2), eps=1e-6, atol=1e-5)


def test_reference_epoch_runs_and_learns():
    g = karate_graph()
    torch.manual_seed(0)
    gcn = O.OracleGCN(34, 4, n_hidden_gcn=16, dropout=0.5)
    opt = torch.optim.Adam(gcn.parameters(), lr=0.02)
    losses = [O.reference_epoch(gcn, g, opt)[0] for _ in range(60)]
    assert losses[-1] < losses[0]
    assert [n for n, _ in gcn.named_parameters()] == ["layers.0.weight", "layers.0.bias",
                                                      "layers.1.weight", "layers.1.bias"]


def test_oracle_reproduces_committed_golden_vectors():
    """tests/golden/gcn_karate.npz (made by oracle/make_golden.py): pins the oracle's numbers -- CSR of A_hat
    bit for bit, logits / loss / gradients to fp32 round-off -- against drift of torch or of this file."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gcn_karate.npz"))
    ei, ea = torch.from_numpy(z["edge_index"]), torch.from_numpy(z["edge_attr"])
    rowptr, col, val, dis, _ = O.csr_from_gcn_norm(ei, ea, 34)
    assert np.array_equal(rowptr.numpy(), z["rowptr"]) and np.array_equal(col.numpy(), z["colidx"])
    assert np.array_equal(val.numpy().view(np.int32), z["val"].view(np.int32))
    assert np.array_equal(dis.numpy().view(np.int32), z["dis"].view(np.int32))
    g = karate_graph()
    W = [torch.from_numpy(z["W1"]).requires_grad_(), torch.from_numpy(z["W2"]).requires_grad_()]
    b = [torch.from_numpy(z["b1"]).requires_grad_(), torch.from_numpy(z["b2"]).requires_grad_()]
    out = O.gcn_forward(g.x, ei, ea, W, b, p=0.5, training=True, drop_masks=[torch.from_numpy(z["keep"])])
    loss = O.masked_cross_entropy(out, torch.from_numpy(z["y"]), torch.from_numpy(z["train_mask"]))
    loss.backward()
    assert rel_err(out, torch.from_numpy(z["logits"])) < 1e-6 and abs(loss.item() - float(z["loss"])) < 1e-6
    for t, k in ((W[0], "gW1"), (b[0], "gb1"), (W[1], "gW2"), (b[1], "gb2")):
        assert rel_err(t.grad, torch.from_numpy(z[k])) < 1e-6
