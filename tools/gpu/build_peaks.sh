#!/bin/bash
# builds tools/gpu/peaks (git-ignored binary; travels to the GPU box with the snapshot)
cd "$(dirname "$0")/../.." && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -o tools/gpu/peaks tools/gpu/peaks.cu -lcuda
