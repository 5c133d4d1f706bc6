mkdir -p gpurun_out
# launch list of one full step of the bench workload (music256), after 3 warm-up steps: 116 launches/step incl. state reset memsets (not kernels)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 345 -c 115 --csv --log-file gpurun_out/launches_r1_final_music256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
# the dominant kernel (fused DWS tensor-core GEMM), full set, 2 launches: a narrow high-rate layer and a wide one
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 120 -c 1 -o gpurun_out/prof_r1_final_dws_wide python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1; echo "ncu full rc=$?"
