mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "upsample_fused and 3-64-32" -x > gpurun_out/sanitizer_up.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/sanitizer_up.log | head -60 | cut -c1-250
