#!/bin/bash
OUT=gpurun_out/r2m
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernel_families.py tests/test_gpu_full_size.py -m gpu -q --tb=short -k "tile or encoder_cfg2" 2>&1 | tail -8 | tee $OUT/pytest.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --option fwd_variant=5 2>>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tile5', d['kernel_ms'])"
# the ncu launch list of the default bench command (op part), and a full capture of the default kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -f -o $OUT/prof_enc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 > $OUT/ncu_enc.log 2>&1
tail -1 $OUT/ncu_enc.log | cut -c1-150
tail -3 $OUT/err.log
