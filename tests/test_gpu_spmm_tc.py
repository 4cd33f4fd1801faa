"""GPU parity of the hybrid propagation (dense blocks on tcgen05 with 3xTF32, the rest gathered) against the gather
kernel, an fp64 evaluation and the oracle-checked trainer."""
import pytest
import torch

from helpers import rel_err
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(cuda, name="small", seed=0, density=0.01):
    from pytextgcn_b200.graph import upload_graph
    from pytextgcn_b200.synthetic import make_graph
    from pytextgcn_b200.tc_plan import build_tc_plan
    g = make_graph(name, seed=seed)
    n = int(g.x.shape[0])
    gr = upload_graph(g.edge_index.to(cuda), g.edge_attr.to(cuda), n)
    tc = build_tc_plan(gr, min_density=density, n_sms=torch.cuda.get_device_properties(cuda).multi_processor_count)
    return g, n, gr, tc


def _fp64(gr, B, bias):
    n = gr.n_nodes
    A = torch.sparse_coo_tensor(torch.stack([gr.row_ids(), gr.colidx.long()]), gr.val.double(), size=(n, n)).coalesce()
    return torch.sparse.mm(A, B.double()) + bias.double()


@pytest.mark.parametrize("F", [64, 100, 200, 256])
@pytest.mark.parametrize("density", [1 / 4096, 0.01, 0.1])
def test_hybrid_matches_gather_and_fp64(cuda, F, density):
    from pytextgcn_b200 import ops
    g, n, gr, tc = _setup(cuda, density=density)
    assert tc is not None and tc.nnz_dense > 0
    torch.manual_seed(F)
    B, bias = torch.randn(n, F, device=cuda), torch.randn(F, device=cuda)
    ref, _ = ops.spmm(gr, B, bias=bias)
    out, _ = ops.spmm_hybrid(tc, B, bias=bias, plan=tc.remainder.plan())
    z64 = _fp64(gr, B, bias)
    assert rel_err(out, ref) < TOL and rel_err(out, z64) < TOL
    out2, _ = ops.spmm_hybrid(tc, B, bias=bias, plan=tc.remainder.plan())
    assert torch.equal(out, out2)                                     # fixed summation order: bitwise repeatable


def test_hybrid_carries_every_epilogue(cuda):
    """bias + Philox dropout, strided operand, and the fused Adam update on the hybrid path vs the gather path."""
    from pytextgcn_b200 import ops
    g, n, gr, tc = _setup(cuda)
    F = 200
    torch.manual_seed(3)
    Bbig = torch.randn(n, F + 56, device=cuda)
    B, bias = Bbig[:, :F], torch.randn(F, device=cuda)
    step = torch.full((1,), 4, dtype=torch.int64, device=cuda)
    kw = dict(bias=bias, drop_mode=ops.DROP_PHILOX, drop_p=0.5, philox_seed=9, philox_offset_dev=step)
    a, _ = ops.spmm(gr, B, F=F, **kw)
    b, _ = ops.spmm_hybrid(tc, B, F=F, plan=tc.remainder.plan(), **kw)
    assert torch.equal(a == 0, b == 0) and rel_err(b, a) < TOL

    def adam_state():
        torch.manual_seed(5)
        return dict(param=torch.randn(n, F, device=cuda), exp_avg=torch.zeros(n, F, device=cuda),
                    exp_avg_sq=torch.zeros(n, F, device=cuda), max_exp_avg_sq=torch.zeros(n, F, device=cuda))
    res = []
    for hybrid in (False, True):
        st = adam_state()
        sd, hyper = torch.zeros(1, dtype=torch.int64, device=cuda), torch.zeros(2, device=cuda)
        ops.adam_prepare(sd, hyper, 0.05)
        if hybrid:
            ops.spmm_hybrid(tc, B, F=F, plan=tc.remainder.plan(), want_out=False, adam=dict(hyper=hyper, **st))
        else:
            ops.spmm(gr, B, F=F, want_out=False, adam=dict(hyper=hyper, **st))
        res.append(st)
    # the first Adam step is lr * g / (|g| + eps): rounding-level differences in tiny gradients are amplified
    assert rel_err(res[1]["exp_avg"], res[0]["exp_avg"]) < TOL and rel_err(res[1]["exp_avg_sq"], res[0]["exp_avg_sq"]) < 2 * TOL
    assert rel_err(res[1]["param"], res[0]["param"]) < 1e-3


def test_trainer_on_tensor_cores_matches_the_reference_epoch(cuda):
    from pytextgcn_b200 import GCN
    from pytextgcn_b200.synthetic import make_graph, SHAPES
    from pytextgcn_b200.trainer import TextGCNTrainer
    shape = SHAPES["small"]
    g = make_graph(shape, seed=1)
    n = int(g.x.shape[0])
    torch.manual_seed(0)
    ref = O.OracleGCN(n, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=0.0)
    mod = GCN(n, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=0.0)
    with torch.no_grad():
        for pd, ps in zip(mod.parameters(), ref.parameters()):
            pd.copy_(ps)
    mod = mod.to(cuda)
    tr = TextGCNTrainer(mod, g.clone().to(cuda), lr=0.01, amsgrad=True, tensor_cores=True, tc_min_density=0.01)
    assert tr.tc is not None and tr.tc.nnz_dense > 0.5 * tr.graph.nnz
    opt = torch.optim.Adam(ref.parameters(), lr=0.01, amsgrad=True)
    for step in range(4):
        out_ref = O.reference_epoch(ref, g, opt)
        out = tr.epoch()
        if step == 0:
            for gbuf, pr in zip(tr.grads, ref.parameters()):
                assert rel_err(gbuf, pr.grad) < TOL
        assert abs(out["loss"] - out_ref[0]) < 1e-5 * max(1, abs(out_ref[0])) * (step + 1)
        assert abs(out["val_loss"] - out_ref[1]) < 1e-4 * max(1, abs(out_ref[1])) * (step + 1)
