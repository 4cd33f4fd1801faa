import os, sys, time, json, torch, torch.distributed as dist
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pytextgcn_b200.dist import DistTextGCNTrainer, shutdown
from pytextgcn_b200.synthetic import make_graph, SHAPES
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "20ng"]
g = make_graph(shape, seed=0)
tr = DistTextGCNTrainer(g, shape.n_classes, shape.hidden, shape.dropout, shape.lr, shape.amsgrad, rank, world, dev, use_cuda_graph=False)
for i in range(4): tr.epoch()
torch.cuda.synchronize(); dist.barrier()
tr.profile = []
K = 10
for i in range(K): tr.epoch()
ph = tr.phase_times_ms()
if rank == 0:
    tot = sum(ph.values())
    print(json.dumps({"world": world, "shape": shape.name, "ms_per_epoch_sum": tot / K, "phases_ms_per_epoch": {k: v / K for k, v in ph.items()}}), flush=True)
shutdown(tr)
