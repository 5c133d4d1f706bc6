mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 27 -c 1 -o gpurun_out/prof_tc_k768 python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_tc2.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/prof_tc_k64 python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_tc3.log 2>&1; echo "ncu rc=$?"
