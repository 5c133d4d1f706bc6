mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tm -s 5 -c 1 -o gpurun_out/prof_tm python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_tm.log 2>&1; echo "ncu rc=$?"
