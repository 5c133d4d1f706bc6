for a in 0 1 2 4 8 9 16 11 27 31; do
  echo "== ablate $a"; HILCODEC_ABLATE=$a timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), 'gemm', round(d['kernel_categories']['pointwise_gemm']['ms_per_step'],2))"
done
