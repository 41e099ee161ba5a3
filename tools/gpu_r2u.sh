#!/bin/bash
# late round 2: final validation + refreshed ncu of the small-Q kernels + sanitizer over the kernel families
OUT=gpurun_out/r2_final2
mkdir -p $OUT
bash tools/gpu_final.sh r2_final2 2>&1 | tail -40
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --model-steps 0"
timeout 300 $B --workload pose_cfg3_t3 --value-dtype bf16 --no-gpu-baseline > $OUT/bench_pose_cfg3_t3_bf16.json 2>>$OUT/err.log
timeout 300 $B --workload petr_cfg1 --fused --no-gpu-baseline > $OUT/bench_petr_cfg1_fused.json 2>>$OUT/err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda -s 8 -c 2 -f -o $OUT/prof_pose python bench.py --workload pose_cfg3 --steps 2 --warmup 3 --launch eager --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 > $OUT/ncu_pose.log 2>&1
tail -1 $OUT/ncu_pose.log | cut -c1-150
bash tools/sanitize_r2.sh r2_final2/san 2>&1 | tail -8
