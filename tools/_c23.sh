cd /root/repo
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_train.py tests/test_gpu_spmm.py -x -q > gpurun_out/c23_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c23_tests.log
tail -15 gpurun_out/c23_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --workload scale --steps 10 --warmup 5 --no-extras --partition words > gpurun_out/c23_scale_words.json 2> gpurun_out/c23_scale_words.err; echo "bench rc=$?"
tail -c 900 gpurun_out/c23_scale_words.json; tail -3 gpurun_out/c23_scale_words.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tools/dist_phases.py scale words > gpurun_out/c23_phases_words.json 2> gpurun_out/c23_phases_words.err; echo "phases rc=$?"
cat gpurun_out/c23_phases_words.json
