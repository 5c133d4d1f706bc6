mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "resblock" > gpurun_out/pytest_rb.log 2>&1; echo "pytest rb rc=$?"; tail -40 gpurun_out/pytest_rb.log | cut -c1-400
