cd /root/repo
timeout 900 python bench.py > gpurun_out/c27_bench_n1.json 2> gpurun_out/c27_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/c27_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/c27_ref.json 2> gpurun_out/c27_ref.err; echo "ref rc=$?"
cat gpurun_out/c27_ref.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c27_smoke.log
