# round 2: last check of the committed tree -- full GPU suite, smoke, one short bench line per workload class
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic']); print({k:(v.get('value'), v.get('ms_per_step'), v.get('ms_per_hop')) for k,v in d['other_workloads'].items()})"
