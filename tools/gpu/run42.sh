mkdir -p gpurun_out
cp hilcodec_b200/libhilcodec_b200.so /tmp/lib_new.so
for v in new old new old; do
  if [ $v = new ]; then cp /tmp/lib_new.so hilcodec_b200/libhilcodec_b200.so; else cp hilcodec_b200/alt/librb_old.so hilcodec_b200/libhilcodec_b200.so; fi
  echo "== RB $v"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_rb_$v.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/bench_rb_$v.json').read()); print(round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernel_categories'].items()}, d['clocks'])"
done
cp /tmp/lib_new.so hilcodec_b200/libhilcodec_b200.so
