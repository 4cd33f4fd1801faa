"""CPU tests of the host-side logic that needs no device: feature decoding (x = I / [I|F]), the Data stand-in,
the chunk-length rule, the synthetic generator's layout and the bench helpers."""
import os
import pickle
import sys

import torch

from oracle import gcn_oracle as O
from pytextgcn_b200.graph import auto_chunk_nnz
from pytextgcn_b200.models import decode_features
from pytextgcn_b200.synthetic import SHAPES, GraphShape, make_graph


def test_decode_identity_features():
    x = O.sparse_identity_features(50)
    info = decode_features(x, n_vocab=20)
    assert info is not None and info.Fdoc is None and info.n_nodes == 50 and info.n_cols == 50
    assert decode_features(x, n_vocab=20) is info                       # cached per tensor


def test_decode_identity_plus_hierarchy_features():
    hf = torch.softmax(torch.randn(30, 4), dim=1)
    hf[3] = 0                                                           # an all-zero feature row is legal
    x = O.sparse_identity_features(50, hf, n_vocab=20)
    info = decode_features(x, n_vocab=20)
    assert info is not None and tuple(info.Fdoc.shape) == (30, 4) and info.n_vocab == 20
    assert torch.equal(info.Fdoc, hf)


def test_decode_rejects_general_matrices():
    idx = torch.tensor([[0, 1, 2, 2], [0, 1, 2, 0]])
    x = torch.sparse_coo_tensor(idx, torch.ones(4), size=(3, 3)).coalesce()       # off-diagonal entry
    assert decode_features(x) is None
    y = torch.sparse_coo_tensor(torch.tensor([[0, 1], [0, 1]]), torch.tensor([1.0, 2.0]), size=(2, 2)).coalesce()
    assert decode_features(y) is None                                   # diagonal but not the identity
    assert decode_features(torch.eye(3).to_sparse().coalesce()) is not None


def test_data_stand_in_behaves_like_pyg_data(tmp_path):
    g = make_graph("tiny", seed=0)
    assert set(["x", "edge_index", "edge_attr", "y", "train_mask", "val_mask", "test_mask", "n_vocab"]) <= set(g.keys)
    assert g.num_nodes == 200 and g.num_edges == g.edge_index.shape[1] and "x" in g and g["y"] is g.y
    g._tgcn_graph = ("device cache", object())                          # never pickled, never moved
    g2 = pickle.loads(pickle.dumps(g))
    assert not hasattr(g2, "_tgcn_graph") and torch.equal(g2.edge_index, g.edge_index)
    assert g.to("cpu") is g
    c = g.clone()
    c.y[0] = 7
    assert g.y[0] != 7


def test_synthetic_graph_has_the_reference_layout():
    shape = GraphShape("t", 300, 200, 4000, 15, 5, 32)
    g = make_graph(shape, seed=3)
    V, n = shape.n_words, shape.n_words + shape.n_docs
    ei, ew = g.edge_index, g.edge_attr
    assert ei.dtype == torch.int64 and not ei.is_contiguous() and ew.dtype == torch.float32
    ww = (ei[0] < V) & (ei[1] < V)
    n_ww = int(ww.sum())
    assert torch.all(ww[:n_ww]) and not ww[n_ww:].any()                  # word-word block first
    a, b = ei[:, :n_ww:2], ei[:, 1:n_ww:2]
    assert torch.equal(a[0], b[1]) and torch.equal(a[1], b[0]) and torch.all(a[0] < a[1])    # (i,j),(j,i) pairs, i < j
    key = a[0] * V + a[1]
    assert torch.all(key[1:] > key[:-1])                                 # upper-triangle row-major order
    rest = ei[:, n_ww:]
    m = rest.shape[1] // 2
    assert torch.all(rest[0, :m] >= V) and torch.all(rest[1, :m] < V)    # (doc+V, word) ...
    assert torch.equal(rest[0, m:], rest[1, :m]) and torch.equal(rest[1, m:], rest[0, :m])   # ... then (word, doc+V)
    d = rest[0, :m]
    assert torch.all(d[1:] >= d[:-1])                                    # doc-major
    tf = torch.zeros(shape.n_docs, dtype=torch.float64).index_add_(0, d - V, ew[n_ww:n_ww + m].double() ** 2)
    assert torch.allclose(tf[tf > 0], torch.ones_like(tf[tf > 0]), atol=1e-5)   # L2-normalised rows like TfidfTransformer
    A = torch.zeros(n, n)
    A[ei[0], ei[1]] = ew
    assert torch.equal(A, A.T) and A.diag().abs().sum() == 0
    assert g.y[:V].sum() == 0 and not (g.train_mask & g.val_mask).any() and not g.train_mask[:V].any()
    assert make_graph(shape, seed=3).edge_attr.equal(ew)                 # seeded


def test_named_shapes_match_the_survey():
    s = SHAPES["20ng"]
    assert (s.n_words, s.n_docs, s.hidden, s.n_classes) == (42757, 18846, 200, 20)
    assert SHAPES["scale"].n_docs == 1_000_000 and SHAPES["scale"].hidden == 256
    assert SHAPES["amazon"].dropout == 0.7 and SHAPES["dbpedia"].amsgrad is False


def test_auto_chunk_rule():
    assert auto_chunk_nnz(21_500_000) == 2048 and auto_chunk_nnz(2_690_000) == 512
    assert auto_chunk_nnz(100) == 256 and auto_chunk_nnz(10 ** 9) == 2048
    vals = [auto_chunk_nnz(n) for n in (10 ** k for k in range(3, 10))]
    assert vals == sorted(vals)


def test_bench_helpers_edge_subsample_and_bytes():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    g = make_graph("small", seed=0)
    gs = bench.edge_subsample(g, 0.3, seed=1)
    n = int(g.x.shape[0])
    assert 0.2 < gs.edge_index.shape[1] / g.edge_index.shape[1] < 0.4
    A = torch.zeros(n, n)
    A[gs.edge_index[0], gs.edge_index[1]] = gs.edge_attr
    assert torch.equal(A, A.T)                                           # both directions kept or dropped together
    assert bench.spmm_alg_bytes(21_510_965, 61_603, 200) == 21_510_965 * 8 + 61_604 * 4 + 2 * 61_603 * 200 * 4 + 800
