mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/bench_speech64.log 2>&1; echo "bench64 rc=$?"; tail -3 gpurun_out/bench_speech64.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_music256.log 2>&1; echo "bench256 rc=$?"; tail -3 gpurun_out/bench_music256.log
