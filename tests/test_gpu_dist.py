"""GPU tests of the row-partitioned trainer.  world_size 1 runs everywhere (it exercises the node
renumbering, padding rows and the sharded kernels' row offsets); the 2-rank NCCL run needs >= 2
GPUs and is skipped on a single-GPU box (the CPU/gloo version of it is tests/test_dist_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import rel_err
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("hier,classes,hidden", [(None, 6, 64), (5, 6, 64), (None, 21, 16), (5, 21, 16)])
def test_world_size_one_matches_oracle(cuda, hier, classes, hidden):
    """classes > hidden exercises the propagate-first order of layer 2 (hidden rows exchanged instead of class-wide ones)."""
    from pytextgcn_b200.dist import DistTextGCNTrainer
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    shape = GraphShape("t", 700, 555, 12000, 20, classes, hidden)
    g = make_graph(shape, seed=3, hierarchy_classes=hier)
    n = int(g.x.shape[0])
    torch.manual_seed(0)
    ref = O.OracleGCN(int(g.x.shape[1]), shape.n_classes, n_hidden_gcn=shape.hidden, dropout=0.0)
    init = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    tr = DistTextGCNTrainer(g, shape.n_classes, shape.hidden, 0.0, 0.01, True, 0, 1, cuda, seed=0, init_weights=init)
    assert tr.propagate_first == (classes > hidden)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01, amsgrad=True)
    for step in range(3):
        out_ref = O.reference_epoch(ref, g, opt)
        tr.train_step()
        assert abs(tr.train_loss() - out_ref[0]) < 1e-5 * max(1, abs(out_ref[0]))
        assert rel_err(tr.part.to_old(tr.g_W1[:tr.part.n_loc]), ref.layers[0].weight.grad[:n]) < 2e-5 * (step + 1)
        if hier:
            assert rel_err(tr.g_W1[tr.part.n_loc:], ref.layers[0].weight.grad[n:]) < 2e-5 * (step + 1)
        assert rel_err(tr.g_W2, ref.layers[1].weight.grad) < 2e-5 * (step + 1)
        assert rel_err(tr.g_b1, ref.layers[0].bias.grad) < 2e-5 * (step + 1)
        assert rel_err(tr.g_b2, ref.layers[1].bias.grad) < 2e-5 * (step + 1)
        tr.eval_step()
        st = tr.epoch_stats()
        assert abs(st["val_loss"] - out_ref[1]) < 1e-4 * max(1, abs(out_ref[1]))
        assert abs(st["acc_val"] - out_ref[3]) < 0.02
    z = tr.logits_old_order()
    ref.eval()
    rows = g.train_mask | g.val_mask | g.test_mask                       # restrict_rows: logits exist on the masked rows
    with torch.no_grad():
        assert rel_err(z[rows.to(z.device)], ref(g)[rows]) < 1e-4
    for k, v in ref.state_dict().items():
        assert rel_err(tr.gathered_parameters()[k], v) < 1e-3, k


def test_dropout_mask_consistent_between_forward_and_backward_on_a_shard(cuda):
    from pytextgcn_b200.dist import DistTextGCNTrainer
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    shape = GraphShape("t", 400, 333, 6000, 20, 6, 32)
    g = make_graph(shape, seed=4)
    tr = DistTextGCNTrainer(g, shape.n_classes, shape.hidden, 0.5, 0.0, False, 0, 1, cuda, seed=5)
    for _ in range(3):
        tr.train_step()
        dropped = tr.H1d == 0
        assert torch.all(tr.dZ1_loc[dropped] == 0)
        assert abs((~dropped).float().mean().item() - 0.5) < 0.02


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_matches_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_renumbered_single_gpu_run_is_the_oracle_of_the_partitioned_run(cuda):
    """world_size 1 (runs on any box): the partitioned trainer with dropout, CUDA graph and the shared hidden
    activation against the single-GPU trainer on the renumbered graph -- the check `bench.py --gpus N` prints."""
    from pytextgcn_b200.dist import parity_against_single_gpu
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    if not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29641")
        torch.distributed.init_process_group("gloo", rank=0, world_size=1)
    try:
        shape = GraphShape("t", 500, 433, 8000, 20, 6, 32, dropout=0.5, amsgrad=True, lr=0.02)
        g = make_graph(shape, seed=6)
        par = parity_against_single_gpu(g, shape, 0, 1, cuda, seed=2, epochs=6, use_cuda_graph=True, keep_w1_grad=False)
        assert par["cuda_graph"]
        # one rank, same kernels, same order of every sum: bit-identical up to the plan's chunk length
        assert par["max_rel_err_loss"] < 1e-5 and par["max_rel_err_W2"] < 1e-4 and par["max_rel_err_W1"] < 1e-4, par
    finally:
        torch.distributed.destroy_process_group()


def test_word_block_trainer_single_rank(cuda):
    """world = 1 exercises the whole word-block data path (split CSR pieces, partial word rows added in the SpMM
    epilogue, restricted class-wide propagations, shared hidden rows, fused Adam): against the oracle's reference epoch
    (dropout off), and -- dropout on, CUDA graph -- against the single-GPU trainer on the renumbered graph."""
    from pytextgcn_b200.dist import parity_against_single_gpu
    from pytextgcn_b200.dist_bipartite import BipartiteTextGCNTrainer
    from pytextgcn_b200.synthetic import make_graph, GraphShape
    shape = GraphShape("t", 700, 1555, 12000, 20, 6, 64)
    g = make_graph(shape, seed=3)
    n = int(g.x.shape[0])
    torch.manual_seed(0)
    ref = O.OracleGCN(n, shape.n_classes, n_hidden_gcn=shape.hidden, dropout=0.0)
    with torch.no_grad():
        for l in ref.layers:
            l.bias.uniform_(-0.1, 0.1)
    init = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    tr = BipartiteTextGCNTrainer(g, shape.n_classes, shape.hidden, 0.0, 0.01, True, 0, 1, cuda, seed=0, init_weights=init)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01, amsgrad=True)
    for step in range(3):
        out_ref = O.reference_epoch(ref, g, opt)
        tr.train_step()
        assert abs(tr.train_loss() - out_ref[0]) < 1e-5 * max(1, abs(out_ref[0]))
        assert rel_err(tr.part.to_old(tr.g_W1), ref.layers[0].weight.grad) < 2e-5 * (step + 1)
        assert rel_err(tr.g_W2, ref.layers[1].weight.grad) < 2e-5 * (step + 1)
        assert rel_err(tr.g_b1, ref.layers[0].bias.grad) < 2e-5 * (step + 1)
        tr.eval_step()
        st = tr.epoch_stats()
        assert abs(st["val_loss"] - out_ref[1]) < 1e-4 * max(1, abs(out_ref[1]))
    for k, v in ref.state_dict().items():
        assert rel_err(tr.gathered_parameters()[k], v) < 1e-3, k
    del tr
    if not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29643")
        torch.distributed.init_process_group("gloo", rank=0, world_size=1)
    try:
        shape = GraphShape("t2", 500, 1433, 8000, 20, 6, 32, dropout=0.5, amsgrad=True, lr=0.02)
        g = make_graph(shape, seed=6)
        par = parity_against_single_gpu(g, shape, 0, 1, cuda, seed=2, epochs=6, partition="words", use_cuda_graph=True,
                                        keep_w1_grad=False)
        assert par["partition"] == "BipartitePartition" and par["cuda_graph"], par
        assert par["max_rel_err_loss"] < 1e-4 and par["max_rel_err_W2"] < 1e-3 and par["max_rel_err_W1"] < 1e-3, par
    finally:
        torch.distributed.destroy_process_group()
