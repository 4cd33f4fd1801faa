cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 10 --warmup 5 --no-extras --workload dbpedia-perlevel > gpurun_out/c30_perlevel_n8.json 2> gpurun_out/c30_perlevel_n8.err; echo "perlevel rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c30_perlevel_n8.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["extra"]["partition"], d["extra"]["cuda_graph"])
PY
