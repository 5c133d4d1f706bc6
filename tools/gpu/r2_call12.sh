mkdir -p gpurun_out
timeout 600 python tools/gpu/trace_gemm.py 256 2> gpurun_out/r2c12_trace.log; echo "trace rc=$?"
