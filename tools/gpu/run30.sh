mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "resblock" > gpurun_out/pytest_rb.log 2>&1; echo "pytest rb rc=$?"; tail -30 gpurun_out/pytest_rb.log | cut -c1-300
for v in "" 1; do
  echo "== HILCODEC_RB_WIDE=$v"; env ${v:+HILCODEC_RB_WIDE=$v} timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_rbw$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_rbw$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, 'launches', d['gpu_launches'], 'checksum', d['e2e']['checksum'])"
done
HILCODEC_RB_WIDE=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:resblock_kernel -c 12 --csv --log-file gpurun_out/launches_rb.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rb.log 2>&1; echo "ncu rc=$?"
