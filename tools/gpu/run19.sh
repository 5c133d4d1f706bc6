mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_codec.py tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "graph or stft or golden or fixture" > gpurun_out/pytest_part.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_part.log | cut -c1-300
timeout 300 python tools/gpu/stream_time.py
