"""GPU parity: graph upload (tgcn_csr_from_coo_gcn_norm) must reproduce the oracle's gcn_norm
CSR BIT FOR BIT -- row pointers, column indices, values, deg^-1/2 -- through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import karate_graph, random_graph
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu


def _check(ei, w, n, cuda, keep_slot=True):
    from pytextgcn_b200.graph import upload_graph
    g = upload_graph(ei.to(cuda) if ei.is_contiguous() else ei.T.contiguous().to(cuda).T, None if w is None else w.to(cuda),
                     n, keep_edge_slot=keep_slot)
    rowptr, col, val, dis, perm = O.csr_from_gcn_norm(ei, w if w is not None else torch.ones(ei.shape[1]), n)
    assert g.nnz == int(rowptr[-1])
    assert torch.equal(g.rowptr.cpu().long(), rowptr)
    assert torch.equal(g.colidx.cpu().long(), col)
    assert torch.equal(g.val.cpu().view(torch.int32), val.view(torch.int32)), "values not bit-exact"
    assert torch.equal(g.dis.cpu().view(torch.int32), dis.view(torch.int32)), "deg^-1/2 not bit-exact"
    if keep_slot:
        # edge_slot maps gcn_norm's edge list (original edges, then loops) to CSR slots
        slot = g.edge_slot.cpu().long()
        ei2, w_hat = O.gcn_norm(ei, w if w is not None else torch.ones(ei.shape[1]), n)
        kept = torch.cat([ei[0] != ei[1], torch.ones(n, dtype=torch.bool)])
        assert torch.all(slot[~kept] == -1)
        s = slot[kept]
        assert torch.equal(g.val.cpu()[s].view(torch.int32), w_hat.view(torch.int32))
        assert torch.equal(g.colidx.cpu().long()[s], ei2[0])
    return g


def test_karate(cuda):
    g = karate_graph()
    _check(g.edge_index, g.edge_attr, 34, cuda)


@pytest.mark.parametrize("case", ["sym", "directed", "loops", "dups", "isolated", "view", "noweight"])
def test_random_graphs(cuda, case):
    n = 500
    kw = dict(sym=dict(), directed=dict(symmetric=False), loops=dict(self_loops=40), dups=dict(duplicates=300),
              isolated=dict(isolated=60), view=dict(transposed_view=True), noweight=dict())[case]
    ei, w = random_graph(n, 20000, seed=11, **kw)
    _check(ei, None if case == "noweight" else w, n, cuda)


def test_hub_rows_long_sequential_sums(cuda):
    # a few hub targets with tens of thousands of in-edges: the sequential fp32 degree sum must
    # still match torch CPU's scatter_add_ order exactly
    rng = np.random.default_rng(0)
    n = 3000
    hubs = rng.integers(0, 5, size=60000)
    others = rng.integers(5, n, size=60000)
    s = np.concatenate([others, hubs, rng.integers(0, n, 20000)])
    d = np.concatenate([hubs, others, rng.integers(0, n, 20000)])
    keep = s != d
    ei = torch.from_numpy(np.stack([s[keep], d[keep]]).astype(np.int64))
    w = torch.from_numpy(rng.uniform(1e-4, 9.0, size=ei.shape[1]).astype(np.float32))
    _check(ei, w, n, cuda)


def test_empty_edge_list(cuda):
    ei = torch.zeros((2, 0), dtype=torch.int64)
    g = _check(ei, torch.zeros(0), 7, cuda)
    assert g.nnz == 7 and torch.all(g.val.cpu() == 1.0)


def test_bad_index_raises(cuda):
    from pytextgcn_b200.graph import upload_graph
    ei = torch.tensor([[0, 1, 9], [1, 0, 2]], device=cuda)
    with pytest.raises(RuntimeError):
        upload_graph(ei, torch.ones(3, device=cuda), 5)


def test_cpu_tensor_raises():
    from pytextgcn_b200.graph import upload_graph
    with pytest.raises(RuntimeError):
        upload_graph(torch.tensor([[0, 1], [1, 0]]), torch.ones(2), 2)


def test_synthetic_textgcn_shape_r8(cuda):
    # R8-shape graph in the reference's edge layout (non-contiguous coo.T view)
    from pytextgcn_b200.synthetic import make_graph
    g = make_graph("r8", seed=0)
    assert not g.edge_index.is_contiguous()
    gc = _check(g.edge_index, g.edge_attr, int(g.x.shape[0]), cuda, keep_slot=False)
    assert gc.is_symmetric()


def test_symmetry_detection(cuda):
    from pytextgcn_b200.graph import upload_graph
    ei, w = random_graph(200, 3000, seed=1, symmetric=False)
    g = upload_graph(ei.to(cuda), w.to(cuda), 200)
    assert not g.is_symmetric()
    t = g.transpose()
    # transpose of transpose has the same entries as the original
    A = torch.zeros(200, 200, dtype=torch.float64)
    rows = g.row_ids().cpu()
    A.index_put_((rows, g.colidx.cpu().long()), g.val.cpu().double(), accumulate=True)
    At = torch.zeros(200, 200, dtype=torch.float64)
    At.index_put_((t.row_ids().cpu(), t.colidx.cpu().long()), t.val.cpu().double(), accumulate=True)
    assert torch.equal(A.T.contiguous(), At)


def test_csr_sidecar_roundtrip(cuda, tmp_path):
    """a cached-CSR sidecar lets a later run skip the sort; it is only accepted for the graph it was
    built from (fingerprint), otherwise the CSR is rebuilt."""
    from pytextgcn_b200.graph import upload_graph, upload_graph_cached, load_csr
    ei, w = random_graph(400, 8000, seed=3, transposed_view=True)
    ei_d, w_d = ei.T.contiguous().to(cuda).T, w.to(cuda)
    path = str(tmp_path / "TGData_123.csr")
    a = upload_graph_cached(ei_d, w_d, 400, path)
    b = upload_graph_cached(ei_d, w_d, 400, path)                     # second call: loaded from the sidecar
    ref = upload_graph(ei_d, w_d, 400)
    for x in (a, b):
        assert torch.equal(x.rowptr, ref.rowptr) and torch.equal(x.colidx, ref.colidx)
        assert torch.equal(x.val.view(torch.int32), ref.val.view(torch.int32)) and torch.equal(x.dis, ref.dis)
    w2 = w_d.clone()
    w2[5] += 1.0
    assert load_csr(path, ei_d, w2, 400) is None                       # different weights -> sidecar refused
    assert load_csr(str(tmp_path / "missing.csr"), ei_d, w_d, 400) is None
