mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log | cut -c1-250
