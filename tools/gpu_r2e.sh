#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_kernel_families.py -m gpu -q --tb=short 2>&1 | tail -60 > $OUT/pytest.log; tail -30 $OUT/pytest.log
