# round 2: config 5 (2048 x 24000 over 8 GPUs) and the 2-GPU point, launched as the driver does
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2s_bench_$n.json 2> gpurun_out/r2s_bench_$n.err
echo "bench N=$n rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_$n.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('gather'), d['clocks'])"
done
timeout 600 python -m pytest tests/test_gpu_sharding_nccl.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
