mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train_graph.py tests/test_gpu_next_rows.py -q --tb=short -p no:cacheprovider > gpurun_out/tg_pytest_new.log 2>&1; echo "new rc=$?"; tail -30 gpurun_out/tg_pytest_new.log | cut -c1-400
timeout 300 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider --deselect tests/test_gpu_train_graph.py --deselect tests/test_gpu_next_rows.py > gpurun_out/tg_pytest_rest.log 2>&1; echo "rest rc=$?"; tail -4 gpurun_out/tg_pytest_rest.log | cut -c1-300
timeout 120 python tools/gpu/stream_time.py > gpurun_out/stream_time.log 2>&1; echo "stream rc=$?"; cat gpurun_out/stream_time.log | tail -6
timeout 120 python __graft_entry__.py --smoke > gpurun_out/tg_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/tg_smoke.log
