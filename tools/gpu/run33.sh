mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "upsample_fused or dwconv" > gpurun_out/pytest_up.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_up.log | cut -c1-250
