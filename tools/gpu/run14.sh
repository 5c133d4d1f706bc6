mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "stft" > gpurun_out/pytest_stft.log 2>&1; echo "stft rc=$?"; tail -12 gpurun_out/pytest_stft.log | cut -c1-250
