"""Host-side logic of the multi-GPU path on CPU: the snake row partition, CSR sharding into the
padded id space, and -- under a real 2-process gloo group -- the exact collective sequence of
DistTextGCNTrainer (all_gather_into_tensor between layers, all_reduce of the small gradients)
with the local SpMMs emulated by torch CPU index ops, checked against the single-process oracle."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import rel_err
from oracle import gcn_oracle as O
from pytextgcn_b200.dist import RowPartition, shard_csr
from pytextgcn_b200.synthetic import make_graph, GraphShape

SHAPE = GraphShape("t", 300, 257, 5000, 14, 5, 16)     # N = 557: not divisible by 2 or 4 -> padding rows


def _global_csr(g):
    n = int(g.x.shape[0])
    rowptr, col, val, dis, _ = O.csr_from_gcn_norm(g.edge_index, g.edge_attr, n)
    return n, rowptr, col, val


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_balances_rows_and_nnz(world):
    g = make_graph(SHAPE, seed=0)
    n, rowptr, col, val = _global_csr(g)
    row_nnz = rowptr[1:] - rowptr[:-1]
    part = RowPartition(row_nnz, world)
    assert part.n_pad == part.n_loc * world and part.n_pad >= n
    # bijection between old ids and the non-padding new ids
    assert torch.equal(torch.sort(part.new_id).values, torch.sort(part.old_id[part.old_id >= 0] * 0 +
                                                                 torch.nonzero(part.old_id >= 0).view(-1)).values)
    assert torch.equal(part.old_id[part.new_id], torch.arange(n))
    nnz = [part.nnz_of(r) for r in range(world)]
    assert sum(nnz) == int(row_nnz.sum())
    assert max(nnz) - min(nnz) <= 2 * int(row_nnz.max())            # snake order: imbalance bounded by ~one heavy row
    x = torch.randn(n, 3)
    assert torch.equal(part.to_old(part.to_new(x)), x)


def _local_spmm(rp, ci, v, B):
    rows = torch.repeat_interleave(torch.arange(rp.numel() - 1), (rp[1:] - rp[:-1]).long())
    out = torch.zeros(rp.numel() - 1, B.shape[1], dtype=B.dtype)
    out.index_add_(0, rows, v.to(B.dtype).view(-1, 1) * B[ci.long()])
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_spmm_equals_global(world):
    g = make_graph(SHAPE, seed=1)
    n, rowptr, col, val = _global_csr(g)
    part = RowPartition(rowptr[1:] - rowptr[:-1], world)
    B = torch.randn(n, 8, dtype=torch.float64)
    ref = _local_spmm(rowptr, col, val, B)
    Bn = part.to_new(B)
    outs = []
    for r in range(world):
        rp, ci, v = shard_csr(rowptr, col, val, part, r)
        assert rp.numel() == part.n_loc + 1 and int(ci.max()) < part.n_pad
        outs.append(_local_spmm(rp, ci, v, Bn))
    assert rel_err(part.to_old(torch.cat(outs)), ref) < 1e-12


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        g = make_graph(SHAPE, seed=2)
        n, rowptr, col, val = _global_csr(g)
        H, C = SHAPE.hidden, SHAPE.n_classes
        part = RowPartition(rowptr[1:] - rowptr[:-1], world)
        rp, ci, v = shard_csr(rowptr, col, val, part, rank)
        nl, lo = part.n_loc, rank * part.n_loc
        W1 = torch.randn(n, H, dtype=torch.float64) * 0.2
        b1 = torch.randn(H, dtype=torch.float64) * 0.1
        W2 = torch.randn(H, C, dtype=torch.float64) * 0.2
        b2 = torch.randn(C, dtype=torch.float64) * 0.1
        y = part.to_new(g.y)[lo:lo + nl]
        tm = part.to_new(g.train_mask, False)[lo:lo + nl]
        n_train = int(g.train_mask.sum())
        # --- the collective sequence of DistTextGCNTrainer.train_step (dropout off) ---
        W1_full = torch.zeros(part.n_pad, H, dtype=torch.float64)
        W1_full[lo:lo + nl] = part.to_new(W1)[lo:lo + nl]                      # only the own shard is valid
        dist.all_gather_into_tensor(W1_full, W1_full[lo:lo + nl].clone())
        H1 = _local_spmm(rp, ci, v, W1_full) + b1
        P_full = torch.zeros(part.n_pad, C, dtype=torch.float64)
        dist.all_gather_into_tensor(P_full, (H1 @ W2).contiguous())
        Z2 = _local_spmm(rp, ci, v, P_full) + b2
        logp = torch.log_softmax(Z2, dim=1)
        dZ2 = torch.zeros_like(Z2)
        rows = torch.nonzero(tm).view(-1)
        dZ2[rows] = torch.exp(logp[rows])
        dZ2[rows, y[rows]] -= 1.0
        dZ2 /= n_train
        loss_part = -logp[rows, y[rows]].sum().view(1)
        dZ2_full = torch.zeros(part.n_pad, C, dtype=torch.float64)
        dist.all_gather_into_tensor(dZ2_full, dZ2.contiguous())
        G2 = _local_spmm(rp, ci, v, dZ2_full)
        small = torch.cat([(G2 @ W2.T).sum(0), (H1.T @ G2).reshape(-1), dZ2.sum(0)])   # db1, dW2, db2
        dist.all_reduce(small)
        dist.all_reduce(loss_part)
        dZ1_full = torch.zeros(part.n_pad, H, dtype=torch.float64)
        dist.all_gather_into_tensor(dZ1_full, (G2 @ W2.T).contiguous())
        gW1_loc = _local_spmm(rp, ci, v, dZ1_full)
        gW1_full = torch.zeros(part.n_pad, H, dtype=torch.float64)
        dist.all_gather_into_tensor(gW1_full, gW1_loc.contiguous())
        if rank == 0:
            # --- single-process oracle ---
            Wr = [W1.clone().requires_grad_(), W2.clone().requires_grad_()]
            br = [b1.clone().requires_grad_(), b2.clone().requires_grad_()]
            z = O.gcn_forward(g.x.double(), g.edge_index, g.edge_attr.double(), Wr, br)
            loss = O.masked_cross_entropy(z, g.y, g.train_mask)
            loss.backward()
            res = dict(loss=abs(loss_part.item() / n_train - loss.item()),
                       gW1=rel_err(part.to_old(gW1_full), Wr[0].grad),
                       db1=rel_err(small[:H], br[0].grad), dW2=rel_err(small[H:H + H * C].view(H, C), Wr[1].grad),
                       db2=rel_err(small[H + H * C:], br[1].grad))
            q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _pf_worker(rank, world, port, q):
    """The collective sequence of DistTextGCNTrainer's propagate-first order (classes > hidden): the exchanged operands
    are the hidden rows and dZ2 W2^T (N x H), the class-wide products stay local."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        shape = GraphShape("t", 300, 257, 5000, 14, 23, 8)          # 23 classes > 8 hidden units
        g = make_graph(shape, seed=2)
        n, rowptr, col, val = _global_csr(g)
        H, C = shape.hidden, shape.n_classes
        part = RowPartition(rowptr[1:] - rowptr[:-1], world)
        rp, ci, v = shard_csr(rowptr, col, val, part, rank)
        nl, lo = part.n_loc, rank * part.n_loc
        W1 = torch.randn(n, H, dtype=torch.float64) * 0.2
        b1 = torch.randn(H, dtype=torch.float64) * 0.1
        W2 = torch.randn(H, C, dtype=torch.float64) * 0.2
        b2 = torch.randn(C, dtype=torch.float64) * 0.1
        y = part.to_new(g.y)[lo:lo + nl]
        tm = part.to_new(g.train_mask, False)[lo:lo + nl]
        n_train = int(g.train_mask.sum())

        def gather(loc):
            full = torch.zeros(part.n_pad, loc.shape[1], dtype=torch.float64)
            dist.all_gather_into_tensor(full, loc.contiguous())
            return full
        H1 = _local_spmm(rp, ci, v, gather(part.to_new(W1)[lo:lo + nl])) + b1          # own rows of A_hat W1 + b1
        U = _local_spmm(rp, ci, v, gather(H1))                                           # A_hat H1 (hidden-wide exchange)
        Z2 = U @ W2 + b2
        logp = torch.log_softmax(Z2, dim=1)
        dZ2 = torch.zeros_like(Z2)
        rows = torch.nonzero(tm).view(-1)
        dZ2[rows] = torch.exp(logp[rows])
        dZ2[rows, y[rows]] -= 1.0
        dZ2 /= n_train
        loss_part = -logp[rows, y[rows]].sum().view(1)
        dZ1 = _local_spmm(rp, ci, v, gather(dZ2 @ W2.T))                                 # A_hat (dZ2 W2^T)
        small = torch.cat([dZ1.sum(0), (U.T @ dZ2).reshape(-1), dZ2.sum(0)])             # db1, dW2 = U^T dZ2, db2
        dist.all_reduce(small)
        dist.all_reduce(loss_part)
        gW1_full = gather(_local_spmm(rp, ci, v, gather(dZ1)))
        if rank == 0:
            Wr = [W1.clone().requires_grad_(), W2.clone().requires_grad_()]
            br = [b1.clone().requires_grad_(), b2.clone().requires_grad_()]
            z = O.gcn_forward(g.x.double(), g.edge_index, g.edge_attr.double(), Wr, br)
            loss = O.masked_cross_entropy(z, g.y, g.train_mask)
            loss.backward()
            q.put(dict(loss=abs(loss_part.item() / n_train - loss.item()),
                       gW1=rel_err(part.to_old(gW1_full), Wr[0].grad),
                       db1=rel_err(small[:H], br[0].grad), dW2=rel_err(small[H:H + H * C].view(H, C), Wr[1].grad),
                       db2=rel_err(small[H + H * C:], br[1].grad)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_propagate_first_sequence_matches_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_pf_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["loss"] < 1e-6
    for k in ("gW1", "db1", "dW2", "db2"):
        assert res[k] < 1e-5, (k, res[k])


def test_two_rank_gloo_collective_sequence_matches_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # fp64 emulation: only the edge values (fp32 A_hat) limit the agreement
    assert res["loss"] < 1e-6
    for k in ("gW1", "db1", "dW2", "db2"):
        assert res[k] < 1e-5, (k, res[k])


# --------------------------------------------------------------------------------------
# word-block (bipartite) partition of dist_bipartite.py
# --------------------------------------------------------------------------------------
from pytextgcn_b200.dist_bipartite import BipartitePartition, shard_bipartite  # noqa: E402


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_bipartite_partition_is_a_bijection_with_words_first(world):
    g = make_graph(SHAPE, seed=0)
    n, rowptr, col, val = _global_csr(g)
    part = BipartitePartition(rowptr[1:] - rowptr[:-1], g.n_vocab, world)
    assert part.n_loc == part.v_loc + part.d_loc and part.n_pad == world * part.n_loc
    assert torch.equal(part.old_id[part.new_id], torch.arange(n))
    loc = part.new_id % part.n_loc
    assert bool((loc[:g.n_vocab] < part.v_loc).all()) and bool((loc[g.n_vocab:] >= part.v_loc).all())
    x = torch.randn(n, 3)
    assert torch.equal(part.to_old(part.to_new(x)), x)
    # words and documents are both balanced to within one row per rank
    for lo, hi in ((0, g.n_vocab), (g.n_vocab, n)):
        per = torch.bincount(part.new_id[lo:hi] // part.n_loc, minlength=world)
        assert int(per.max() - per.min()) <= 1


@pytest.mark.parametrize("world", [1, 2, 3])
def test_bipartite_sharded_spmm_equals_global(world):
    """main_r @ [all words ; own rows] + sum_s (q_s @ docs of s)[words of r] == rows of A_hat X owned by r."""
    g = make_graph(SHAPE, seed=1)
    n, rowptr, col, val = _global_csr(g)
    part = BipartitePartition(rowptr[1:] - rowptr[:-1], g.n_vocab, world)
    nl, vl, vp = part.n_loc, part.v_loc, part.v_pad
    B = torch.randn(n, 8, dtype=torch.float64)
    ref = _local_spmm(rowptr, col, val, B)
    Bn = part.to_new(B).view(world, nl, 8)
    words = Bn[:, :vl].reshape(vp, 8)                       # what the all-gather of the word block delivers
    shards = [shard_bipartite(rowptr, col, val, part, r) for r in range(world)]
    assert sum(int(m[1].numel() + q[1].numel()) for m, q in shards) == int(col.numel())      # every entry exactly once
    partial = [_local_spmm(*q, Bn[r, vl:]) for r, (m, q) in enumerate(shards)]               # [vp, 8] per rank
    outs = []
    for r, (m, q) in enumerate(shards):
        assert m[0].numel() == nl + 1 and int(m[1].max()) < vp + nl
        o = _local_spmm(*m, torch.cat([words, Bn[r]]))
        for s in range(world):                              # the all-to-all: slot s = rank s's partial rows of my words
            o[:vl] += partial[s][r * vl:(r + 1) * vl]
        outs.append(o)
    assert rel_err(part.to_old(torch.cat(outs)), ref) < 1e-12


def test_bipartite_rejects_document_document_edges():
    g = make_graph(SHAPE, seed=1)
    n, rowptr, col, val = _global_csr(g)
    part = BipartitePartition(rowptr[1:] - rowptr[:-1], g.n_vocab - 5, 2)      # pretend 5 words are documents
    with pytest.raises(NotImplementedError):
        shard_bipartite(rowptr, col, val, part, 0)


def _bip_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = make_graph(SHAPE, seed=2)
        n, rowptr, col, val = _global_csr(g)
        part = BipartitePartition(rowptr[1:] - rowptr[:-1], g.n_vocab, world)
        nl, vl, vp, lo = part.n_loc, part.v_loc, part.v_pad, rank * part.n_loc
        main, qq = shard_bipartite(rowptr, col, val, part, rank)
        torch.manual_seed(0)
        X = torch.randn(n, 6, dtype=torch.float64)
        X_loc = part.to_new(X)[lo:lo + nl]
        # the collective sequence of BipartiteTextGCNTrainer._propagate
        OP = torch.zeros(vp + nl, 6, dtype=torch.float64)       # [gathered word block ; this rank's own rows]
        OP[vp:] = X_loc
        dist.all_gather_into_tensor(OP[:vp], OP[vp:vp + vl].contiguous())
        Q = _local_spmm(*qq, OP[vp + vl:]).contiguous()
        SL = torch.zeros(world, vl, 6, dtype=torch.float64)
        recv = list(SL.unbind(0))
        for d in range(world):      # all_to_all_single on NCCL; gloo has no all-to-all, so one gather per destination
            dist.gather(Q.view(world, vl, 6)[d].contiguous(), recv if rank == d else None, dst=d)
        out = _local_spmm(*main, OP)
        for s in range(world):
            out[:vl] += recv[s]
        full = torch.zeros(part.n_pad, 6, dtype=torch.float64)
        dist.all_gather_into_tensor(full, out.contiguous())
        if rank == 0:
            q.put(rel_err(part.to_old(full), _local_spmm(rowptr, col, val, X)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_word_block_exchange_matches_global():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bip_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res < 1e-12


def test_partition_choice_rule():
    """auto = word-block when x = I and the word block + partial word rows (2 V rows) are at most 3/4 of the N rows the
    row partition moves; [I | F] inputs and word-heavy graphs keep the row partition."""
    from types import SimpleNamespace
    from pytextgcn_b200.dist import choose_partition

    def g(n_vocab, n_docs, extra_cols=0):
        n = n_vocab + n_docs
        return SimpleNamespace(x=torch.empty((n, n + extra_cols), device="meta"), n_vocab=n_vocab)
    assert choose_partition(g(42757, 18846)) == "row"            # 20NG: words outnumber documents
    assert choose_partition(g(200_000, 1_000_000)) == "words"    # scale configuration
    assert choose_partition(g(20_000, 50_000)) == "words"        # Amazon-shape
    assert choose_partition(g(10_000, 337_739, extra_cols=9)) == "row"   # per-level input [I | onehot(parent)]
    assert choose_partition(g(10_000, 337_739)) == "words"
    assert choose_partition(g(200_000, 1_000_000), "row") == "row" and choose_partition(g(42757, 18846), "words") == "words"
