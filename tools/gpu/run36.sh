mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "upsample_fused" > gpurun_out/pytest_up.log 2>&1; rc=$?; echo "up rc=$rc"; tail -12 gpurun_out/pytest_up.log | cut -c1-250
if [ $rc -ne 0 ]; then exit 0; fi
export HILCODEC_FUSE_UPSAMPLE=1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_up1.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_up1.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, 'launches', d['gpu_launches'], 'checksum', d['e2e']['checksum'])"
timeout 600 ncu --kernel-name-base mangled -k regex:gemm_h_kernelILb0ELi[2458]E --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -c 4 --csv --log-file gpurun_out/launches_up.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_up.log 2>&1; echo "ncu rc=$?"
