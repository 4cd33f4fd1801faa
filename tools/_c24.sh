cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 > gpurun_out/c24_bench_n8.json 2> gpurun_out/c24_bench_n8.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c24_bench_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 tools/dist_phases.py scale words > gpurun_out/c24_phases_words.json 2> gpurun_out/c24_phases_words.err; echo "phases rc=$?"
cat gpurun_out/c24_phases_words.json
