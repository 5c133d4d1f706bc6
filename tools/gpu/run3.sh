mkdir -p gpurun_out
timeout 300 python tools/gpu/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; tail -15 gpurun_out/tc_check.log
