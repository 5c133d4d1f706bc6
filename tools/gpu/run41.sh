mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider -k "resblock" > gpurun_out/pytest_rb.log 2>&1; rc=$?; echo "rb rc=$rc"; tail -5 gpurun_out/pytest_rb.log | cut -c1-250
for i in 1 2; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r41.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_r41.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['e2e']['checksum'])"; done
