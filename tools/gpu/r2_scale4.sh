mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2s_bench_4.json 2> gpurun_out/r2s_bench_4.err
echo "bench N=4 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_4.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('gather'), d['clocks'])"
