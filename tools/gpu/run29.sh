mkdir -p gpurun_out
# resblock launches in a speech64 step: s0 x2, s1 x2, s2 x2 (BN64), u2 x3 (BN64), u3 x3 -> the 10th..12th are u3 (C=96)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resblock_kernel -s 45 -c 1 -o gpurun_out/prof_rb_u3 python bench.py --steps 1 --warmup 3 --workload speech64 --no-cpu-baseline > gpurun_out/ncu_full_rb.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
