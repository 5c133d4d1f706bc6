mkdir -p gpurun_out
# 1. full GPU parity suite
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
# 2. default bench (music256) incl. cpu baseline
timeout 600 python bench.py > gpurun_out/bench_music256.json 2> gpurun_out/bench_music256.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_music256.json
# 3. launch list of one step (+ a bit) after 3 warm-up steps
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/launches_r1b_music256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
# 4. full-set capture of two gemm_h launches (speech64 is enough: same kernel, smaller batch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_h_kernel -s 100 -c 2 -o gpurun_out/prof_r1b_gemm_h python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
