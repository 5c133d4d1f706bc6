mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/verify_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/verify_pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/verify_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/verify_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/verify_bench.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['bound'], round(d['roofline']['frac'],3), d['roofline']['traffic'], d['cpu_baseline']['value'], d['clocks'])"
