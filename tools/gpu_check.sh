#!/bin/bash
# One GPU-box visit: parity tests, smoke, micro-benchmark, bench, ncu launch list.
# Usage (under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
lscpu | head -20 > $OUT/cpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== microbench" ; timeout 120 ./tools/microbench_red 2>&1 | tee $OUT/microbench_red.log
echo "== bench (default flags)" ; timeout 600 python bench.py 2>$OUT/bench.err > $OUT/bench.json; cut -c1-400 $OUT/bench.json
tail -5 $OUT/bench.err
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1 stress_cfg5; do
  timeout 300 python bench.py --steps 100 --warmup 10 --workload $wl 2>>$OUT/bench.err > $OUT/bench_$wl.json
done
timeout 300 python bench.py --steps 100 --warmup 10 --value-dtype bf16 --no-cpu-baseline --no-e2e 2>>$OUT/bench.err > $OUT/bench_bf16.json
timeout 300 python bench.py --steps 100 --warmup 10 --value-dtype bf16 --workload pose_cfg3 --no-cpu-baseline --no-e2e 2>>$OUT/bench.err > $OUT/bench_pose_cfg3_bf16.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json; cut -c1-300 $OUT/bench_reference.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda -c 40 --csv --log-file $OUT/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log | cut -c1-300
if [ -n "$NCU_FULL" ]; then
  echo "== ncu full"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -f -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log | cut -c1-300
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -f -o $OUT/prof_pose python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload pose_cfg3 > $OUT/ncu_full_pose.log 2>&1
  ls -la $OUT/*.ncu-rep
fi
echo "== modules"; timeout 300 python tools/bench_modules.py 2>&1 | tee $OUT/bench_modules.jsonl
echo done
