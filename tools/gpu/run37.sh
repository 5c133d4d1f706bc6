mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r37.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_r37.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, 'launches', d['gpu_launches'], 'checksum', d['e2e']['checksum'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -s 219 -c 80 --csv --log-file gpurun_out/launches_r37.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r37.log 2>&1; echo "ncu rc=$?"
