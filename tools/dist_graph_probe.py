import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pytextgcn_b200.dist import DistTextGCNTrainer
from pytextgcn_b200.synthetic import make_graph, SHAPES
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "small"]
g = make_graph(shape, seed=0)
tr = DistTextGCNTrainer(g, shape.n_classes, shape.hidden, shape.dropout, shape.lr, shape.amsgrad, rank, world, dev, use_cuda_graph=True)
t0 = time.time()
for i in range(8):
    tr.epoch()
    torch.cuda.synchronize()
    if rank == 0: print(f"epoch {i} done t={time.time()-t0:.2f} graph={tr._graph is not None} err={tr.graph_error}", flush=True)
st = tr.epoch_stats()
if rank == 0: print("stats", st, flush=True)
torch.cuda.synchronize()
t0 = time.time()
for i in range(20): tr.epoch()
torch.cuda.synchronize()
if rank == 0: print(f"20 epochs {1e3*(time.time()-t0)/20:.3f} ms/epoch", flush=True)
from pytextgcn_b200.dist import shutdown; shutdown(tr); print('exit', rank, flush=True)
