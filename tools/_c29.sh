cd /root/repo
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/c29_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c29_tests.log
tail -12 gpurun_out/c29_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 5 --no-extras --workload dbpedia-perlevel > gpurun_out/c29_perlevel.json 2> gpurun_out/c29_perlevel.err; echo "perlevel rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/c29_perlevel.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["extra"]["partition"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 tools/dist_phases.py dbpedia row > gpurun_out/c29_phases_l3.json 2> gpurun_out/c29_phases_l3.err; echo "phases rc=$?"; cat gpurun_out/c29_phases_l3.json
