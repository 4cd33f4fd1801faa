"""GPU parity of the training-side kernels: masked NLL, fused Adam/AMSGrad, the dense backward
pass, and the fused TextGCNTrainer (CUDA-graph epochs) against the oracle's reference epoch."""
import pytest
import torch

from helpers import rel_err
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("C", [3, 20, 64, 219])
def test_masked_nll_and_gradient(cuda, C):
    from pytextgcn_b200 import ops
    torch.manual_seed(C)
    n = 3000
    z = (torch.randn(n, C) * 3).requires_grad_()
    y = torch.randint(0, C, (n,))
    mask = torch.rand(n) > 0.6
    y_masked = y.clone()
    y_masked[~mask] = -1                               # never read where mask == 0
    loss_ref = O.masked_cross_entropy(z, y, mask)
    loss_ref.backward()
    Cp = ops.pad4(C)
    zd = torch.zeros(n, Cp, device=cuda)
    zd[:, :C] = z.detach().to(cuda)
    r = ops.masked_nll(zd, C, y_masked.to(cuda), mask.to(cuda), int(mask.sum()), want_grad=True, want_pred=True,
                       want_correct=True)
    assert abs(r["loss"][0].item() - loss_ref.item()) < TOL * abs(loss_ref.item())
    assert int(r["loss"][1].item()) == int(mask.sum())
    assert rel_err(r["dZ"][:, :C], z.grad) < TOL
    assert torch.all(r["dZ"][:, C:] == 0)
    pred_ref = z.detach().numpy().argmax(axis=1)
    assert (r["pred"].cpu().numpy()[mask.numpy()] == pred_ref[mask.numpy()]).all()
    assert int(r["correct"].item()) == int((torch.from_numpy(pred_ref)[mask] == y[mask]).sum())


@pytest.mark.parametrize("amsgrad", [False, True])
def test_adam_matches_torch(cuda, amsgrad):
    from pytextgcn_b200 import ops
    torch.manual_seed(0)
    n = 10007                                           # odd length: vector body + scalar tail
    p_ref = torch.randn(n).requires_grad_()
    opt = torch.optim.Adam([p_ref], lr=0.05, amsgrad=amsgrad)
    p = p_ref.detach().clone().to(cuda)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    vmax = torch.zeros_like(p) if amsgrad else None
    step_dev = torch.zeros(1, dtype=torch.int64, device=cuda)
    for step in range(1, 6):
        grad = torch.randn(n) * (0.1 if step % 2 else 2.0)
        p_ref.grad = grad.clone()
        opt.step()
        ops.increment_step(step_dev)
        ops.adam_step(p, grad.to(cuda), m, v, vmax, lr=0.05, amsgrad=amsgrad, step_dev=step_dev)
        assert rel_err(p, p_ref) < 2e-6
    assert int(step_dev.item()) == 5


@pytest.mark.parametrize("H,C,act", [(200, 20, 0), (100, 64, 0), (32, 219, 0), (256, 20, 1), (64, 6, 0), (256, 220, 0), (32, 300, 0)])
def test_dense_backward(cuda, H, C, act):
    from pytextgcn_b200 import ops
    torch.manual_seed(H + C)
    n, p = 777, 0.5
    G2 = torch.randn(n, C)
    keep = torch.rand(n, H) > p
    z1 = torch.randn(n, H)
    h_pre = torch.relu(z1) if act else z1
    H1d = h_pre * keep * (1 / (1 - p))
    W2 = torch.randn(H, C) * 0.1
    dZ2 = torch.randn(n, C)
    # oracle: autograd through hd = dropout(act(z1)); out = hd @ W2, with upstream grad G2
    z1r = z1.clone().requires_grad_()
    W2r = W2.clone().requires_grad_()
    hd = (torch.relu(z1r) if act else z1r) * keep * (1 / (1 - p))
    (hd @ W2r).backward(G2)
    Cp = ops.pad4(C)
    G2d = torch.zeros(n, Cp, device=cuda); G2d[:, :C] = G2.to(cuda)
    dZ2d = torch.zeros(n, Cp, device=cuda); dZ2d[:, :C] = dZ2.to(cuda)
    r = ops.dense_bwd(G2d, H1d.to(cuda), W2.to(cuda), dZ2d, H=H, n_classes=C, act=act, drop_mode=ops.DROP_MASK,
                      drop_p=p, keep_mask=keep.to(torch.uint8).to(cuda))
    assert rel_err(r["dW2"], W2r.grad) < TOL
    assert rel_err(r["dZ1"][:, :H], z1r.grad) < TOL
    assert rel_err(r["db_hidden"], z1r.grad.sum(0)) < 5e-5
    assert rel_err(r["db_out"], dZ2.sum(0)) < 5e-5


def _make_pair(cuda, shape_name="small", p=0.0, amsgrad=True, hier=None, relu=False):
    from pytextgcn_b200 import GCN
    from pytextgcn_b200.synthetic import make_graph, SHAPES
    from pytextgcn_b200.trainer import TextGCNTrainer
    shape = SHAPES[shape_name]
    g = make_graph(shape, seed=1, hierarchy_classes=hier)
    n, in_ch = int(g.x.shape[0]), int(g.x.shape[1])
    torch.manual_seed(0)
    ref = O.OracleGCN(in_ch, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=p, relu=relu)
    mod = GCN(in_ch, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=p, apply_activation=relu)
    with torch.no_grad():
        for pd, ps in zip(mod.parameters(), ref.parameters()):
            pd.copy_(ps)
    mod = mod.to(cuda)
    gd = g.clone().to(cuda)
    return g, gd, ref, mod, shape


@pytest.mark.parametrize("graph_mode", [False, True])
@pytest.mark.parametrize("hier", [None, 5])
def test_trainer_gradients_and_first_steps_match_reference_epoch(cuda, graph_mode, hier):
    from pytextgcn_b200.trainer import TextGCNTrainer
    g, gd, ref, mod, shape = _make_pair(cuda, p=0.0, hier=hier)
    tr = TextGCNTrainer(mod, gd, lr=0.01, amsgrad=True, use_cuda_graph=graph_mode)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01, amsgrad=True)
    for step in range(5):                                 # crosses the eager warm-up -> graph replay boundary
        out_ref = O.reference_epoch(ref, g, opt)
        out = tr.epoch()
        # gradients of this step (trainer keeps them in static buffers)
        for gbuf, pr in zip(tr.grads, ref.parameters()):
            assert rel_err(gbuf, pr.grad) < 2e-5 * (step + 1), f"step {step}"
        assert abs(out["loss"] - out_ref[0]) < 1e-5 * max(1, abs(out_ref[0])) * (step + 1)
        assert abs(out["val_loss"] - out_ref[1]) < 1e-4 * max(1, abs(out_ref[1])) * (step + 1)
        assert abs(out["acc_train"] - out_ref[2]) < 0.01 and abs(out["acc_val"] - out_ref[3]) < 0.02
    assert tr.step == 5
    # parameters were updated IN PLACE: the wrapped module is the trained model (state_dict/th.save keep working)
    for pm, pr in zip(mod.parameters(), ref.parameters()):
        assert rel_err(pm, pr) < 1e-3


def test_trainer_with_dropout_learns_and_is_reproducible(cuda):
    from pytextgcn_b200.trainer import TextGCNTrainer
    runs = []
    for _ in range(2):
        g, gd, ref, mod, shape = _make_pair(cuda, p=0.5)
        tr = TextGCNTrainer(mod, gd, lr=0.05, amsgrad=True, seed=7)
        runs.append([tr.epoch()["loss"] for _ in range(12)])
    assert runs[0] == runs[1]                              # deterministic kernels + counter-based Philox
    assert runs[0][-1] < runs[0][0]


def test_trainer_dropout_mask_differs_per_step_and_matches_backward(cuda):
    from pytextgcn_b200.trainer import TextGCNTrainer
    g, gd, ref, mod, shape = _make_pair(cuda, p=0.5)
    tr = TextGCNTrainer(mod, gd, lr=0.0, amsgrad=False, seed=3)   # lr 0: weights frozen, only the mask moves
    masks = []
    for _ in range(4):
        tr.train_step()
        masks.append((tr.H1d != 0).clone())
        # backward used the same mask: dZ1 is zero exactly where the forward dropped
        assert torch.all(tr.dZ1[~masks[-1]] == 0)
    assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[2], masks[3])
    assert abs(masks[0].float().mean().item() - 0.5) < 0.01


def test_trainer_relu_variant(cuda):
    from pytextgcn_b200.trainer import TextGCNTrainer
    g, gd, ref, mod, shape = _make_pair(cuda, p=0.0, relu=True)
    tr = TextGCNTrainer(mod, gd, lr=0.01, amsgrad=False)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01)
    for _ in range(3):
        out_ref = O.reference_epoch(ref, g, opt)
        out = tr.epoch()
        assert abs(out["loss"] - out_ref[0]) < 5e-5 * max(1, abs(out_ref[0]))


@pytest.mark.parametrize("hier", [None, 5])
def test_collapsed_eval_matches_layered_eval_and_oracle(cuda, hier):
    from pytextgcn_b200.trainer import TextGCNTrainer
    g, gd, ref, mod, shape = _make_pair(cuda, p=0.5, hier=hier)
    with torch.no_grad():
        for l in ref.layers:
            l.bias.uniform_(-0.3, 0.3)
        for pd, ps in zip(mod.parameters(), ref.parameters()):
            pd.copy_(ps)
    tr = TextGCNTrainer(mod, gd, lr=0.01, amsgrad=True)
    ref.eval()
    with torch.no_grad():
        z_ref = ref(g)
    out = {}
    for mode in ("layered", "collapsed"):
        tr.set_eval_mode(mode)
        for _ in range(4):                       # eager, eager, capture, replay
            r = tr.eval_step(full=True)
        out[mode] = (r["logits"].clone(), r["val_loss"].clone(), int(r["correct_val"].item()))
        assert rel_err(out[mode][0], z_ref) < 1e-5, mode
    assert rel_err(out["collapsed"][0], out["layered"][0]) < 1e-5
    assert abs(out["collapsed"][1][0].item() - out["layered"][1][0].item()) < 1e-5
    assert abs(out["collapsed"][2] - out["layered"][2]) <= 1


def test_collapsed_eval_refused_with_activation(cuda):
    from pytextgcn_b200.trainer import TextGCNTrainer
    g, gd, ref, mod, shape = _make_pair(cuda, p=0.0, relu=True)
    with pytest.raises(ValueError):
        TextGCNTrainer(mod, gd, eval_mode="collapsed")


def test_trainer_many_classes_narrow_hidden_with_hierarchy(cuda):
    """DBPedia per-level form (perlevel_dbpedia.py:140-141,186): hidden 32, 219 classes, x = [I | onehot(70)]:
    un-fused projection, narrow-hidden dense backward, per-warp hierarchy-gradient kernel."""
    from pytextgcn_b200 import GCN
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    from pytextgcn_b200.trainer import TextGCNTrainer
    shape = GraphShape("t", 600, 900, 8000, 12, 219, 32)
    g = make_graph(shape, seed=2, hierarchy_classes=70)
    n, in_ch = int(g.x.shape[0]), int(g.x.shape[1])
    torch.manual_seed(0)
    ref = O.OracleGCN(in_ch, 219, n_hidden_gcn=32, dropout=0.0)
    mod = GCN(in_ch, 219, n_hidden_gcn=32, dropout=0.0)
    with torch.no_grad():
        for pd, ps in zip(mod.parameters(), ref.parameters()):
            pd.copy_(ps)
    mod = mod.to(cuda)
    tr = TextGCNTrainer(mod, g.clone().to(cuda), lr=0.01, amsgrad=False)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01)
    for step in range(3):
        out_ref = O.reference_epoch(ref, g, opt)
        out = tr.epoch()
        for gbuf, pr in zip(tr.grads, ref.parameters()):
            assert rel_err(gbuf, pr.grad) < 2e-5 * (step + 1)
        assert abs(out["loss"] - out_ref[0]) < 1e-5 * max(1, abs(out_ref[0])) * (step + 1)


@pytest.mark.parametrize("amsgrad", [False, True])
def test_adam_small_multi_tensor_matches_torch(cuda, amsgrad):
    from pytextgcn_b200 import ops
    torch.manual_seed(1)
    shapes = [(200,), (200, 20), (20,)]
    refs = [torch.randn(*s).requires_grad_() for s in shapes]
    opt = torch.optim.Adam(refs, lr=0.05, amsgrad=amsgrad)
    ps = [r.detach().clone().to(cuda) for r in refs]
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    xs = [torch.zeros_like(p) for p in ps]
    step_dev = torch.zeros(1, dtype=torch.int64, device=cuda)
    for step in range(4):
        gs = [torch.randn(*s) for s in shapes]
        for r, g in zip(refs, gs):
            r.grad = g.clone()
        opt.step()
        ops.increment_step(step_dev)
        ops.adam_step_small(ps, [g.to(cuda) for g in gs], ms, vs, xs, lr=0.05, amsgrad=amsgrad, step_dev=step_dev)
        for p, r in zip(ps, refs):
            assert rel_err(p, r) < 2e-6


@pytest.mark.parametrize("mode", ["philox", "mask"])
def test_dropout_apply_equals_the_spmm_epilogue(cuda, mode):
    """tgcn_dropout_apply(A B + bias) must be bit-identical to the SpMM with the dropout epilogue."""
    from pytextgcn_b200 import ops
    from pytextgcn_b200.graph import upload_graph
    from pytextgcn_b200.synthetic import make_graph
    g = make_graph("small", seed=3)
    n, F = int(g.x.shape[0]), 200
    gr = upload_graph(g.edge_index.to(cuda), g.edge_attr.to(cuda), n)
    torch.manual_seed(1)
    B, bias = torch.randn(n, F, device=cuda), torch.randn(F, device=cuda)
    step = torch.full((1,), 5, dtype=torch.int64, device=cuda)
    if mode == "philox":
        kw = dict(drop_mode=ops.DROP_PHILOX, drop_p=0.5, philox_seed=11, philox_offset=3, philox_offset_dev=step)
    else:
        kw = dict(drop_mode=ops.DROP_MASK, drop_p=0.3, keep_mask=(torch.rand(n, F, device=cuda) > 0.3).to(torch.uint8))
    fused, _ = ops.spmm(gr, B, bias=bias, **kw)
    pre, _ = ops.spmm(gr, B, bias=bias)
    two_step = ops.dropout_apply(pre, **kw)
    assert torch.equal(fused, two_step)
    assert 0.2 < (fused == 0).float().mean().item() < 0.6


@pytest.mark.parametrize("graph_mode", [False, True])
@pytest.mark.parametrize("relu", [False, True])
def test_shared_hidden_activation_is_bit_identical_and_saves_a_wide_spmm(cuda, graph_mode, relu, monkeypatch):
    """share_h1: the eval forward's A_hat (X W1) + b1 serves the next training forward (same W1/b1,
    flat_amazon.py:100-110).  Losses, parameters and logits must be bit-identical to recomputing it."""
    from pytextgcn_b200 import ops
    from pytextgcn_b200.trainer import TextGCNTrainer
    hist, wide_calls = {}, {}
    real_spmm = ops.spmm
    for share in (False, True):
        g, gd, ref, mod, shape = _make_pair(cuda, p=0.5, relu=relu)
        tr = TextGCNTrainer(mod, gd, lr=0.05, amsgrad=True, seed=7, share_h1=share, use_cuda_graph=graph_mode)
        count = [0]

        def counting(graph, B, *a, F=None, **k):
            if F == shape.hidden:
                count[0] += 1
            return real_spmm(graph, B, *a, F=F, **k)
        monkeypatch.setattr(ops, "spmm", counting)
        out = []
        for ep in range(6):
            if ep == 3:
                with torch.no_grad():                       # an in-place edit of W1 must invalidate the shared activation
                    mod.layers[0].weight.mul_(1.0)
            r = tr.epoch()
            out.append((r["loss"], r["val_loss"], r["acc_val"]))
        monkeypatch.setattr(ops, "spmm", real_spmm)
        hist[share] = (out, [p_.detach().clone() for p_ in mod.parameters()], tr.logits.clone())
        wide_calls[share] = count[0]
    assert hist[True][0] == hist[False][0]
    for a, b in zip(hist[True][1], hist[False][1]):
        assert torch.equal(a, b)
    assert torch.equal(hist[True][2], hist[False][2])
    if not graph_mode:
        # 3 wide propagations per epoch without sharing; with it 2, plus one for the very first step and one after the edit
        assert wide_calls[False] == 18 and wide_calls[True] == 14



@pytest.mark.parametrize("graph_mode", [False, True])
def test_restricted_class_wide_propagations_change_nothing_that_is_read(cuda, graph_mode):
    """restrict_rows: logits only for the rows a mask selects, backward only over the columns where dZ2 != 0.
    Losses, accuracies, every gradient and the parameters must match the unrestricted run (the skipped terms are
    exact zeros); the masked rows of `logits`/`pred` must equal the full eval bit for bit; changing a mask rebuilds
    the lists."""
    from pytextgcn_b200.trainer import TextGCNTrainer
    res = {}
    for restrict in (False, True):
        g, gd, ref, mod, shape = _make_pair(cuda, p=0.5)
        tr = TextGCNTrainer(mod, gd, lr=0.05, amsgrad=True, seed=11, restrict_rows=restrict, use_cuda_graph=graph_mode)
        hist = [tuple(tr.epoch().values()) for _ in range(5)]
        rows = gd.train_mask | gd.val_mask | gd.test_mask
        masked_logits, masked_pred = tr.logits[rows].clone(), tr.pred[rows].clone()
        full = tr.eval_step(full=True)["logits"].clone()
        assert torch.equal(full[rows], masked_logits)
        # move 50 validation rows into the training set: counts, divisor and work lists must follow
        tm, vm = gd.train_mask.clone(), gd.val_mask.clone()
        idx = torch.nonzero(vm).view(-1)[:50]
        tm[idx], vm[idx] = True, False
        tr.set_masks(gd.y, tm, vm, gd.test_mask)
        hist += [tuple(tr.epoch().values()) for _ in range(4)]
        res[restrict] = (hist, [g_.clone() for g_ in tr.grads], [p_.detach().clone() for p_ in mod.parameters()],
                         masked_logits, masked_pred)
    # The restricted backward matrix has shorter rows, hence other chunk boundaries for the hub rows: the same terms
    # are added in another order, so the two runs agree to fp32 rounding (amplified a little by 9 Adam steps), not bitwise.
    for a, b in zip(res[True][0], res[False][0]):
        assert all(abs(x - y) <= 2e-6 * max(1.0, abs(y)) for x, y in zip(a[:2], b[:2])) and abs(a[2] - b[2]) < 0.01 and abs(a[3] - b[3]) < 0.02
    for a, b in zip(res[True][1], res[False][1]):
        assert rel_err(a, b) < 2e-5
    for a, b in zip(res[True][2], res[False][2]):
        assert rel_err(a, b) < 1e-4
    assert rel_err(res[True][3], res[False][3]) < 1e-5 and (res[True][4] != res[False][4]).float().mean().item() < 0.01


def test_colsum_and_biased_projection(cuda):
    from pytextgcn_b200 import ops
    torch.manual_seed(0)
    X = torch.randn(5003, 36, device=cuda)
    assert rel_err(ops.colsum(X[:, :32], F=32), X[:, :32].double().sum(0)) < 1e-6
    W, b = torch.randn(32, 219, device=cuda), torch.randn(219, device=cuda)
    P = ops.project(X, W, K=32, bias=b)
    assert rel_err(P[:, :219], X[:, :32].double() @ W.double() + b.double()) < 1e-6 and bool((P[:, 219:] == 0).all())


@pytest.mark.parametrize("restrict", [False, True])
def test_propagate_first_layer2_matches_the_reference_order(cuda, restrict):
    """More classes than hidden units: Z2 = (A_hat H1d) W2 + b2 (and its backward) against the oracle, which keeps the
    reference's order A_hat (H1d W2) + b2 -- same linear maps, so logits / loss / gradients agree to fp32 rounding."""
    from pytextgcn_b200 import GCN
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    from pytextgcn_b200.trainer import TextGCNTrainer
    shape = GraphShape("t", 500, 700, 7000, 12, 70, 32)
    g = make_graph(shape, seed=4, hierarchy_classes=9)
    in_ch = int(g.x.shape[1])
    torch.manual_seed(0)
    ref = O.OracleGCN(in_ch, 70, n_hidden_gcn=32, dropout=0.0)
    with torch.no_grad():
        for l in ref.layers:
            l.bias.uniform_(-0.2, 0.2)
    mod = GCN(in_ch, 70, n_hidden_gcn=32, dropout=0.0)
    with torch.no_grad():
        for pd, ps in zip(mod.parameters(), ref.parameters()):
            pd.copy_(ps)
    mod = mod.to(cuda)
    tr = TextGCNTrainer(mod, g.clone().to(cuda), lr=0.01, amsgrad=False, restrict_rows=restrict)
    assert tr.propagate_first
    opt = torch.optim.Adam(ref.parameters(), lr=0.01)
    for step in range(4):
        out_ref = O.reference_epoch(ref, g, opt)
        out = tr.epoch()
        if step == 0:
            for name, gbuf, pr in zip(("W1", "b1", "W2", "b2"), tr.grads, ref.parameters()):
                assert rel_err(gbuf, pr.grad) < TOL, name
        assert abs(out["loss"] - out_ref[0]) < 1e-5 * max(1, abs(out_ref[0])) * (step + 1)
        assert abs(out["val_loss"] - out_ref[1]) < 1e-4 * max(1, abs(out_ref[1])) * (step + 1)
    ref.eval()
    with torch.no_grad():
        z_ref = ref(g)
    z = tr.eval_step(full=True)["logits"]
    assert rel_err(z, z_ref) < 1e-4          # after 4 Adam steps on both sides


def test_propagate_first_dropout_mask_is_the_forward_mask(cuda):
    from pytextgcn_b200.trainer import TextGCNTrainer
    from pytextgcn_b200 import GCN
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    shape = GraphShape("t", 300, 400, 4000, 12, 70, 32)
    g = make_graph(shape, seed=5).to(cuda)
    torch.manual_seed(1)
    mod = GCN(int(g.x.shape[1]), 70, n_hidden_gcn=32, dropout=0.5).to(cuda)
    tr = TextGCNTrainer(mod, g, lr=0.0, amsgrad=False, seed=3, restrict_rows=False)
    assert tr.propagate_first
    for _ in range(3):
        tr.train_step()
        dropped = tr.H1d == 0
        assert torch.all(tr.dZ1[dropped] == 0) and abs((~dropped).float().mean().item() - 0.5) < 0.02


@pytest.mark.parametrize("K,M", [(200, 20), (100, 64), (32, 219), (219, 32), (64, 6), (256, 8)])
def test_row_per_lane_projection_with_fused_dropout(cuda, K, M):
    """tgcn_project_ex: P = dropout(X) W + b in one pass, dropout(X) written out: equal to tgcn_dropout_apply (bitwise)
    followed by an fp64 product; also without dropout, with column tiles (M > 32) and K not a multiple of 4 / 32."""
    from pytextgcn_b200 import ops
    torch.manual_seed(K + M)
    n = 3001
    Kp = ops.pad4(K)
    X = torch.zeros(n, Kp, device=cuda)
    X[:, :K] = torch.randn(n, K, device=cuda)
    W, b = torch.randn(K, M, device=cuda) * 0.2, torch.randn(M, device=cuda)
    P = ops.project(X, W, K=K, bias=b)
    assert rel_err(P[:, :M], X[:, :K].double() @ W.double() + b.double()) < 2e-6 and bool((P[:, M:] == 0).all())
    if K % 4 == 0:
        step = torch.full((1,), 3, dtype=torch.int64, device=cuda)
        kw = dict(drop_mode=ops.DROP_PHILOX, drop_p=0.5, philox_seed=5, philox_offset_dev=step)
        Xd_ref = ops.dropout_apply(X[:, :K], **kw) if Kp == K else None
        Xd = torch.zeros(n, Kp, device=cuda)
        P2 = ops.project(X, W, K=K, dropped_out=Xd, **kw)
        assert torch.equal(Xd[:, :K], Xd_ref)
        assert rel_err(P2[:, :M], Xd_ref.double() @ W.double()) < 2e-6
        keep = (torch.rand(n, K, device=cuda) > 0.3).to(torch.uint8)
        P3 = ops.project(X, W, K=K, drop_mode=ops.DROP_MASK, drop_p=0.3, keep_mask=keep)
        assert rel_err(P3[:, :M], (X[:, :K].double() * keep.double() / 0.7) @ W.double()) < 2e-6


def test_masked_nll_second_accuracy_count(cuda):
    """One pass gives the validation loss / accuracy AND the training-row accuracy (mask2)."""
    from pytextgcn_b200 import ops
    torch.manual_seed(2)
    n, C = 5000, 20
    z = torch.randn(n, C)
    y = torch.randint(0, C, (n,))
    u = torch.rand(n)
    val, train = u < 0.2, (u >= 0.2) & (u < 0.7)
    y_dev = y.clone(); y_dev[~(val | train)] = -1
    zd = z.to(cuda)
    c1, c2 = torch.zeros(1, dtype=torch.int32, device=cuda), torch.zeros(1, dtype=torch.int32, device=cuda)
    r = ops.masked_nll(zd, C, y_dev.to(cuda), val.to(cuda), int(val.sum()), want_grad=False, want_pred=True, correct=c1,
                       mask2=train.to(cuda), correct2=c2)
    pred = z.argmax(1)
    assert int(c1.item()) == int((pred[val] == y[val]).sum()) and int(c2.item()) == int((pred[train] == y[train]).sum())
    assert abs(r["loss"][0].item() - torch.nn.functional.cross_entropy(z[val], y[val]).item()) < 1e-5
    assert torch.equal(r["pred"].cpu().long(), pred)
