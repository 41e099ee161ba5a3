#!/bin/bash
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 300 python tools/exp_footprint.py 2>&1 | tee $OUT/exp_footprint.txt
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_reference_kernel.py tests/test_gpu_kernel_families.py -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --model-steps 0 > $OUT/enc.json 2>$OUT/enc.err; python -c "
import json; d=json.load(open('$OUT/enc.json')); print(d['kernel_ms']); print(d['gpu_baseline'])"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --workload pose_cfg3 > $OUT/pose.json 2>$OUT/pose.err; python -c "
import json; d=json.load(open('$OUT/pose.json')); print(d['kernel_ms']); print(d['gpu_baseline'])"
tail -3 $OUT/*.err
