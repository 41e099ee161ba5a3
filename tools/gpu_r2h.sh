#!/bin/bash
OUT=gpurun_out/r2h
mkdir -p $OUT
(cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_tma_red microbench_tma_red.cu && timeout 200 ./microbench_tma_red) 2>&1 | tee $OUT/microbench_tma_red.txt
