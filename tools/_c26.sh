cd /root/repo
timeout 600 python tools/tc_check.py 20ng 200 0.05 3 > gpurun_out/c26_tc_check.log 2>&1; rc=$?; echo "tc_check rc=$rc"; tail -3 gpurun_out/c26_tc_check.log
if [ $rc -eq 0 ]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_tc_pack|k_tc_mma|k_spmm' -c 7 -o gpurun_out/r02_final_prop -f python tools/tc_check.py 20ng 200 0.05 3 > gpurun_out/c26_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c26_ncu.log
fi
