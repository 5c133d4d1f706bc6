mkdir -p gpurun_out
timeout 300 python tools/gpu/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "rc=$?"; tail -50 gpurun_out/tc_probe.log
