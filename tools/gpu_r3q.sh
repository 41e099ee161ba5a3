#!/bin/bash
# final ncu evidence of the default kernels: launch list of the default op command + full capture of encoder fwd / bwd
OUT=gpurun_out/r3q
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -f -o $OUT/prof_enc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 > $OUT/ncu_enc.log 2>&1
tail -1 $OUT/ncu_enc.log | cut -c1-150
wc -l $OUT/launches.csv
