"""Parity at BASELINE.json's full size (20NG-shape: 61,603 nodes, 2.15e7 non-zeros, hidden 200, 20
classes): gcn_norm's CSR bit for bit, one whole training step + eval forward element for element against
the oracle (logits, loss, all gradients <= 1e-5), and size-independent properties on top: sampled rows
recomputed in fp64 on the host, the adjoint identity of the symmetric operator, determinism, and the
epoch statistics of a few fused training steps."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(cuda):
    from pytextgcn_b200.graph import upload_graph
    from pytextgcn_b200.synthetic import make_graph
    g = make_graph("20ng", seed=0)
    n = int(g.x.shape[0])
    ei = g.edge_index.T.contiguous().to(cuda).T                    # the reference's non-contiguous coo.T view
    ea = g.edge_attr.to(cuda)
    csr = upload_graph(ei, ea, n)
    return g, n, csr, ei, ea


def test_csr_bit_exact_at_full_size(big):
    g, n, csr, _, _ = big
    rowptr, col, val, dis, _ = O.csr_from_gcn_norm(g.edge_index, g.edge_attr, n)
    assert csr.nnz == int(rowptr[-1]) == g.edge_index.shape[1] + n
    assert torch.equal(csr.rowptr.cpu().long(), rowptr)
    assert torch.equal(csr.colidx.cpu().long(), col)
    assert torch.equal(csr.val.cpu().view(torch.int32), val.view(torch.int32))
    assert torch.equal(csr.dis.cpu().view(torch.int32), dis.view(torch.int32))
    assert csr.is_symmetric()


@pytest.mark.parametrize("F", [200, 20])
def test_sampled_rows_against_fp64_host(big, cuda, F):
    from pytextgcn_b200 import ops
    g, n, csr, _, _ = big
    torch.manual_seed(F)
    B = torch.randn(n, F)
    bias = torch.randn(F)
    out, _ = ops.spmm(csr, B.to(cuda), bias=bias.to(cuda))
    rp, ci, v = csr.rowptr.cpu().long(), csr.colidx.cpu().long(), csr.val.cpu().double()
    lens = rp[1:] - rp[:-1]
    rows = torch.cat([torch.argsort(lens, descending=True)[:8],                  # the hub rows (split chunks + fix-up)
                      torch.argsort(lens)[:8], torch.randint(0, n, (240,))])
    ref = torch.stack([(v[rp[r]:rp[r + 1]].view(-1, 1) * B[ci[rp[r]:rp[r + 1]]].double()).sum(0) + bias.double()
                       for r in rows.tolist()])
    got = out[rows.to(cuda)].cpu().double()
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5


def test_adjoint_identity_and_determinism(big, cuda):
    # A_hat is symmetric: <u, A v> == <A u, v>; and the kernels have no atomics: bitwise repeatable
    from pytextgcn_b200 import ops
    g, n, csr, _, _ = big
    torch.manual_seed(1)
    u, v = torch.randn(n, 200, device=cuda), torch.randn(n, 200, device=cuda)
    Av, _ = ops.spmm(csr, v)
    Au, _ = ops.spmm(csr, u)
    a, b = (u.double() * Av.double()).sum().item(), (Au.double() * v.double()).sum().item()
    assert abs(a - b) <= 1e-6 * max(abs(a), abs(b))
    Av2, _ = ops.spmm(csr, v)
    assert torch.equal(Av, Av2)


def test_fused_training_epochs_at_full_size(big, cuda):
    from pytextgcn_b200 import GCN
    from pytextgcn_b200.trainer import TextGCNTrainer
    g, n, csr, ei, ea = big
    gd = g.clone()
    gd.edge_index, gd.edge_attr = ei, ea
    gd = gd.to(cuda)
    losses = []
    for rep in range(2):
        torch.manual_seed(0)
        gcn = GCN(n, 20, n_hidden_gcn=200, dropout=0.5).to(cuda)
        tr = TextGCNTrainer(gcn, gd, lr=0.05, amsgrad=True, seed=3, graph=csr)
        losses.append([tr.epoch() for _ in range(6)])
    first, last = losses[0][0], losses[0][-1]
    assert abs(first["loss"] - np.log(20)) < 0.05                   # Glorot init, 20 balanced classes
    assert last["loss"] < first["loss"] and last["acc_train"] > 0.2
    assert losses[0] == losses[1]                                   # identical run-to-run
    # layered and re-associated eval forward agree at full size too
    tr.set_eval_mode("layered")
    z1 = tr.eval_step(full=True)["logits"].clone()
    tr.set_eval_mode("collapsed")
    z2 = tr.eval_step(full=True)["logits"].clone()
    assert rel_err(z2, z1) < 1e-5


def test_full_train_step_and_eval_against_the_oracle_at_full_size(big, cuda):
    """One whole training step (forward with an explicit dropout keep-mask, masked cross-entropy, backward)
    and the eval forward on the BASELINE 20NG-shape graph, compared element for element with
    oracle.gcn_oracle (the restated GCNConv path, flat_amazon.py:99-110) and with an fp64 evaluation of the same
    formulas: logits, loss and all four gradients within 1e-5 relative (max-norm) of fp64, within 2e-5 of the fp32
    oracle (whose own rounding at this size is measured).  The oracle materialises the (E+N) x 200 message tensors
    (~17 GB each), so this needs ~40 GB of host memory and ~1 min of CPU time."""
    from pytextgcn_b200 import GCN
    g, n, csr, ei, ea = big
    H, C, p = 200, 20, 0.5
    torch.manual_seed(0)
    ref = O.OracleGCN(n, C, n_hidden_gcn=H, dropout=p)
    with torch.no_grad():
        for l in ref.layers:
            l.bias.uniform_(-0.1, 0.1)                      # non-zero biases: the epilogue's bias add is exercised
    keep = torch.rand(n, H) > p
    mod = GCN(n, C, n_hidden_gcn=H, dropout=p)
    with torch.no_grad():
        for pd, ps in zip(mod.parameters(), ref.parameters()):
            pd.copy_(ps)
    mod = mod.to(cuda)
    gd = g.clone()
    gd.edge_index, gd.edge_attr = ei, ea
    gd = gd.to(cuda)

    # ---- CUDA path: the drop-in module, autograd through the C ABI ----
    mod.train()
    mod.drop_mask_override = [keep.to(cuda)]
    z = mod(gd)
    loss = torch.nn.functional.cross_entropy(z[gd.train_mask], gd.y[gd.train_mask])
    loss.backward()
    z_train, loss_train = z.detach().cpu(), float(loss.item())
    grads = [p_.grad.detach().cpu() for p_ in mod.parameters()]
    mod.eval()
    with torch.no_grad():
        z_eval = mod(gd).cpu()
    del z, loss
    torch.cuda.empty_cache()

    # ---- fp64 evaluation of the same formulas (torch sparse CSR on the device, values = gcn_norm's fp32 A_hat):
    #      the yardstick for "whose rounding is it" at this size ----
    rowptr, col, val, _, _ = O.csr_from_gcn_norm(g.edge_index, g.edge_attr, n)
    rows = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    A64 = torch.sparse_coo_tensor(torch.stack([rows, col]).to(cuda), val.double().to(cuda), size=(n, n)).coalesce()
    P64 = [p_.detach().double().to(cuda).requires_grad_() for p_ in ref.parameters()]
    keep64 = keep.to(cuda).double() * (1.0 / (1.0 - p))

    def fwd64(train):
        h = torch.sparse.mm(A64, P64[0]) + P64[1]
        if train:
            h = h * keep64
        return torch.sparse.mm(A64, h @ P64[2]) + P64[3]
    z64 = fwd64(True)
    l64 = torch.nn.functional.cross_entropy(z64[gd.train_mask], gd.y[gd.train_mask])
    l64.backward()
    g64 = [p_.grad.cpu() for p_ in P64]
    z64_train, l64 = z64.detach().cpu(), float(l64.item())
    with torch.no_grad():
        z64_eval = fwd64(False).cpu()
    del A64, z64, keep64
    torch.cuda.empty_cache()
    # the CUDA path against fp64: logits, loss and every gradient within the 1e-5 bar
    assert rel_err(z_train, z64_train) < 1e-5 and rel_err(z_eval, z64_eval) < 1e-5
    assert abs(loss_train - l64) < 1e-5 * abs(l64)
    for name, gm, g6 in zip(("W1", "b1", "W2", "b2"), grads, g64):
        assert rel_err(gm, g6) < 1e-5, name

    # ---- the oracle (fp32, CPU): sequential index_add over rows of up to 14 k terms carries ~1e-5 of rounding of its
    #      own at this size (its distance to fp64 is asserted to be of that order), so two correct fp32 summation
    #      orders can be 2e-5 apart; the CUDA path must be at least as close to fp64 as the oracle is ----
    ref.train()
    zr = ref(g, drop_masks=[keep])
    lr = O.masked_cross_entropy(zr, g.y, g.train_mask)
    lr.backward()
    e_oracle = rel_err(zr.detach(), z64_train)
    assert e_oracle < 2e-5
    assert rel_err(z_train, z64_train) <= max(e_oracle, 2e-6)
    assert rel_err(z_train, zr.detach()) < 2e-5
    assert abs(loss_train - lr.item()) < 1e-5 * abs(lr.item())
    for name, gm, pr in zip(("W1", "b1", "W2", "b2"), grads, ref.parameters()):
        assert rel_err(gm, pr.grad) < 2e-5, name
    del zr, lr
    ref.eval()
    with torch.no_grad():
        zr = ref(g)
    assert rel_err(z_eval, zr) < 2e-5
