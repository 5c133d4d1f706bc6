# round 2, call 5 (2 GPUs): full parity suite incl. the 2-rank NCCL gather test and the cluster RVQ; streaming A/B of the
# cluster RVQ; 2-GPU bench line with the gather timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|config 2:|config 3:|edge cases:|range guard:|NCCL|^E  |FAILED" gpurun_out/r2c5_pytest.log | cut -c1-900 | tail -30
for v in 1 0; do
  for wl in stream1 stream64; do
    HILCODEC_RVQ_CLUSTER=$v timeout 300 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c5_${wl}_cluster$v.json 2> gpurun_out/r2c5_${wl}_cluster$v.err
    echo "$wl rvq_cluster=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c5_${wl}_cluster$v.json').read().strip().splitlines()[-1]); print(d['ms_per_hop'], d['e2e']['ms_per_hop'], d['gpu_launches_per_hop'], {k: round(v['ms'],3) for k,v in d['kernel_categories_per_hop'].items()})"
  done
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c5_bench_2gpu.json 2> gpurun_out/r2c5_bench_2gpu.err
echo "bench 2gpu rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c5_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('gather'))"
tail -2 gpurun_out/r2c5_bench_2gpu.err
