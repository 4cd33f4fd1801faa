cd /root/repo
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 10 --warmup 5 --no-extras "$@" > gpurun_out/c28_$name.json 2> gpurun_out/c28_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c28_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["ms_per_step"], d["extra"]["partition"], d["extra"]["exchange"])
except Exception as e: print("$name", "ERR", e)
PY
}
run amazon_row --workload amazon --partition row
run amazon_words --workload amazon --partition words
run perlevel_row --workload dbpedia-perlevel --partition row
run perlevel_auto --workload dbpedia-perlevel --partition auto
run r8_row --workload r8 --partition row
