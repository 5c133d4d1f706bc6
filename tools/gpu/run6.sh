mkdir -p gpurun_out
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section SchedulerStats --clock-control none -k regex:gemm_tc -c 55 -o gpurun_out/prof_r1_gemm_tc_all python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_tc.log | cut -c1-300
