cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 4 --no-extras > gpurun_out/c34_20ng_n4.json 2> gpurun_out/c34_20ng_n4.err; echo "20ng rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29782 bench.py --gpus 4 --workload scale --steps 10 --warmup 5 --no-extras > gpurun_out/c34_scale_n4.json 2> gpurun_out/c34_scale_n4.err; echo "scale rc=$?"
python - <<PY
import json
for f in ("c34_20ng_n4","c34_scale_n4"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["extra"]["partition"], d["extra"]["exchange"])
PY
