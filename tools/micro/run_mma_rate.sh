#!/bin/bash
# builds (if needed) and runs the tcgen05.mma rate micro-benchmark
cd "$(dirname "$0")"
[ -x mma_rate ] || /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o mma_rate mma_rate.cu -lcuda
./mma_rate
