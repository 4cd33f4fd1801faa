"""Per-phase CUDA-event timing of an epoch of the N-rank trainer (eager mode; the shipped path replays a CUDA graph).
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/dist_phases.py <shape> [auto|row|words]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytextgcn_b200.dist import make_dist_trainer, shutdown  # noqa: E402
from pytextgcn_b200.synthetic import make_graph, SHAPES  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "20ng"]
partition = sys.argv[2] if len(sys.argv) > 2 else "auto"
g = make_graph(shape, seed=0)
tr = make_dist_trainer(g, shape, rank, world, dev, partition, use_cuda_graph=False, keep_w1_grad=False)
for i in range(4):
    tr.epoch()
torch.cuda.synchronize()
dist.barrier()
tr.profile = []
K = 10
for i in range(K):
    tr.epoch()
ph = tr.phase_times_ms()
if rank == 0:
    tot = sum(ph.values())
    print(json.dumps({"world": world, "shape": shape.name, "partition": type(tr.part).__name__, "ms_per_epoch_sum": tot / K,
                      "phases_ms_per_epoch": {k: v / K for k, v in ph.items()}}), flush=True)
shutdown(tr)
