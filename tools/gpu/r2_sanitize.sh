# round 2: compute-sanitizer memcheck over the operator tests that exercise this round's new kernels / paths
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=line -p no:cacheprovider -k "test_pointwise or test_dws_block or test_stft_logmag or test_rvq_tensor_core or test_downsample" > gpurun_out/r2_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r2_sanitize_memcheck.log | head -12
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_codec.py -m gpu -q -x --tb=line -p no:cacheprovider -k "many_streams or graph_streaming or golden_clip" > gpurun_out/r2_sanitize_memcheck_codec.log 2>&1
echo "memcheck codec rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r2_sanitize_memcheck_codec.log | head -12
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=line -p no:cacheprovider -k "test_rvq_tensor_core or test_dws_block or test_pointwise" > gpurun_out/r2_sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitize_racecheck.log | head -12
