#!/bin/bash
# builds tools/gpu/peaks (git-ignored binary; travels to the GPU box with the snapshot)
cd "$(dirname "$0")/../.." && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -o tools/gpu/peaks tools/gpu/peaks.cu -lcuda
# tools/gpu/tma_flat_test: the two tensor maps of the flat tiles in isolation (argv[1] = output swizzle 0 none / 1 128B / 2 32B)
nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -o tools/gpu/tma_flat_test tools/gpu/tma_flat_test.cu -lcuda
