#!/bin/bash
# First GPU call of the next round, in one gpurun invocation (about 15 minutes of box time):
#   1. default GPU suite + the staged-kernel parity tests (TGCN_TEST_STAGED=1), each under its own timeout;
#   2. A/B of the wide propagation: tgcn_spmm vs tgcn_spmm_staged over the plan/launch shapes of tools/ab_spmm.py;
#   3. bench.py with the default kernel and with TGCN_SPMM_STAGED=1 (no CPU arm: it is timed separately);
#   4. one `ncu --set full` capture of the staged kernel (only if step 2 ran clean) for profiles/.
# Usage:  gpurun --timeout 1500 -- 'bash tools/round2_gpu.sh'
# Every step writes into gpurun_out/r02_*; a failing step does not stop the later ones.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_default.log 2>&1
echo "pytest default rc=$?" | tee gpurun_out/r02_status.txt
TGCN_TEST_STAGED=1 timeout 420 python -m pytest tests/test_gpu_spmm_staged.py -q -m gpu > gpurun_out/r02_pytest_staged.log 2>&1
rc_staged=$?
echo "pytest staged rc=$rc_staged" | tee -a gpurun_out/r02_status.txt
timeout 420 python tools/ab_spmm.py 20ng 20 > gpurun_out/r02_ab_spmm_20ng.jsonl 2> gpurun_out/r02_ab_spmm_20ng.err
echo "ab_spmm rc=$?" | tee -a gpurun_out/r02_status.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench default rc=$?" | tee -a gpurun_out/r02_status.txt
if [ "$rc_staged" = "0" ]; then
  TGCN_SPMM_STAGED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_staged.json 2> gpurun_out/r02_bench_staged.err
  echo "bench staged rc=$?" | tee -a gpurun_out/r02_status.txt
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spmm_staged -c 2 \
      -o gpurun_out/r02_prof_spmm_staged -f python tools/ab_spmm.py 20ng 1 --only staged:28,2,64,4,0 > gpurun_out/r02_ncu_staged.log 2>&1
  echo "ncu staged rc=$?" | tee -a gpurun_out/r02_status.txt
fi
# 5. build variants of the gather kernel (same sources, macros flipped): 32-bit gather addressing; (col,val) as one
#    broadcast 8-byte load instead of two shuffles (the L1 data pipe is the saturated unit); all switches incl. exact
#    lanes-per-row for class-wide operands (3.2 instead of 6.4 instructions per non-zero in the narrow loop):
#    parity tests against the oracle with the variant library loaded, then the same bench
for v in addr32 cvpack noalloc all; do
  make -C pytextgcn_b200/csrc variant-$v > gpurun_out/r02_build_$v.log 2>&1 || { echo "build $v failed" | tee -a gpurun_out/r02_status.txt; continue; }
  export TGCN_B200_LIB=$PWD/pytextgcn_b200/lib/libtextgcn_b200_$v.so TGCN_SPMM_CVPACK=1
  timeout 420 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_model.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02_pytest_$v.log 2>&1
  echo "pytest $v rc=$?" | tee -a gpurun_out/r02_status.txt
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err
  echo "bench $v rc=$?" | tee -a gpurun_out/r02_status.txt
  unset TGCN_B200_LIB TGCN_SPMM_CVPACK
done
# 6. 128-byte row pitch for the gathered hidden-wide operands (W1, dZ1), default library and the all-switches variant
TGCN_ROW_ALIGN=1 timeout 300 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02_pytest_rowalign.log 2>&1
echo "pytest rowalign rc=$?" | tee -a gpurun_out/r02_status.txt
TGCN_ROW_ALIGN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_rowalign.json 2> gpurun_out/r02_bench_rowalign.err
echo "bench rowalign rc=$?" | tee -a gpurun_out/r02_status.txt
TGCN_ROW_ALIGN=2 TGCN_SPMM_CVPACK=1 TGCN_B200_LIB=$PWD/pytextgcn_b200/lib/libtextgcn_b200_all.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline \
  > gpurun_out/r02_bench_all_rowalign.json 2> gpurun_out/r02_bench_all_rowalign.err
echo "bench all+rowalign rc=$?" | tee -a gpurun_out/r02_status.txt
cat gpurun_out/r02_status.txt
tail -3 gpurun_out/r02_pytest_default.log gpurun_out/r02_pytest_staged.log
cat gpurun_out/r02_ab_spmm_20ng.jsonl
