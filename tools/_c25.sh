cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c25_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c25_tests.log
tail -4 gpurun_out/c25_tests.log
timeout 900 python bench.py > gpurun_out/c25_bench_n1.json 2> gpurun_out/c25_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/c25_bench_n1.json
timeout 600 python bench.py --workload dbpedia-perlevel --no-cpu-baseline --no-extras > gpurun_out/c25_perlevel.json 2> gpurun_out/c25_perlevel.err; echo "perlevel rc=$?"
timeout 300 python bench.py --steps 3 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c25_short.json 2> gpurun_out/c25_short.err; rc=$?; echo "short rc=$rc"
if [ $rc -eq 0 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c25_launches.csv python bench.py --steps 3 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c25_ncu.log 2>&1; echo "ncu rc=$?"
fi
