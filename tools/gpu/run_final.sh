mkdir -p gpurun_out
# 1. the driver's GPU test tier
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/final_pytest_gpu.log | cut -c1-300
# 2. smoke
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
# 3. default bench (music256, with the CPU baseline) and speech64
timeout 600 python bench.py > gpurun_out/final_bench_music256.json 2> gpurun_out/final_bench_music256.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench_music256.json
timeout 300 python bench.py --workload speech64 --no-cpu-baseline > gpurun_out/final_bench_speech64.json 2>/dev/null; echo "bench speech rc=$?"; cut -c1-300 gpurun_out/final_bench_speech64.json
# 4. launch list of one step after 3 warm-up steps (73 launches per step)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -s 219 -c 80 --csv --log-file gpurun_out/launches_final_music256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final_list.log 2>&1; echo "ncu list rc=$?"
# 5. full-set captures: a wide fused-DWS launch (decoder stage 1, C = 384) and a ResBlock launch (decoder stage 3), speech64 batch
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_h_kernelILb1ELi0E -s 92 -c 1 -o gpurun_out/prof_final_dws_u1 python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_final_full1.log 2>&1; echo "ncu full dws rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resblock_kernel -s 25 -c 1 -o gpurun_out/prof_final_rb_u3 python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_final_full2.log 2>&1; echo "ncu full rb rc=$?"
ls -la gpurun_out | tail -12
