mkdir -p gpurun_out
export HILCODEC_FUSE_UPSAMPLE=1
timeout 600 ncu --kernel-name-base mangled -k regex:gemm_h_kernelILb0ELi[24]E --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -c 4 --csv --log-file gpurun_out/launches_up24.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_up24.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/launches_up24.csv | cut -d, -f5,13,15 | cut -c1-200 | head -30
