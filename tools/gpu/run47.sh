mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_h_kernelILb1ELi0E -s 98 -c 2 -o gpurun_out/prof_dws_u2 python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_u2.log 2>&1; echo "ncu rc=$?"
