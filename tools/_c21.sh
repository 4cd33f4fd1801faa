cd /root/repo
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/c21_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c21_tests.log
tail -5 gpurun_out/c21_tests.log
for part in row words; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --workload scale --steps 10 --warmup 5 --no-extras --partition $part > gpurun_out/c21_scale_$part.json 2> gpurun_out/c21_scale_$part.err; echo "bench $part rc=$?"
tail -c 600 gpurun_out/c21_scale_$part.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tools/dist_phases.py scale $part > gpurun_out/c21_phases_$part.json 2> gpurun_out/c21_phases_$part.err; echo "phases $part rc=$?"
cat gpurun_out/c21_phases_$part.json
done
