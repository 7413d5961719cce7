#!/bin/bash
# One gpurun call: GPU tests, the judged bench (both arms) and the ncu launch list of the bench
# command.  (tools/final_capture.sh: the full ncu captures, exported to CSV on the box.)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_final_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_tests.txt
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_final_reference_arm.json 2>> gpurun_out/r2_bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras \
  --no-cpu-baseline --no-ref-gpu > gpurun_out/r2_launches_bench.log 2>&1
cat gpurun_out/r2_final_tests.txt
head -c 400 gpurun_out/r2_bench_final.json
