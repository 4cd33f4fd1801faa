cd /root/repo
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_train.py -x -q > gpurun_out/c31_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c31_tests.log
tail -6 gpurun_out/c31_tests.log
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 2 --steps 10 --warmup 5 --no-extras "$@" > gpurun_out/c31_$name.json 2> gpurun_out/c31_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c31_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["extra"]["partition"], d["extra"]["exchange"], d["extra"]["kernels_per_epoch"])
except Exception as e: print("$name", "ERR", e)
PY
}
run scale_words --workload scale --partition words
run amazon_words --workload amazon --partition words
