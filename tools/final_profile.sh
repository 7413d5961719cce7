#!/bin/bash
# One gpurun call: the judged bench (both arms), the ncu launch list of the bench command and
# full captures of the dominant kernels (exported to CSV on the box: the .ncu-rep files are too
# large to travel back).  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_final_smi.txt
if [ "$1" != "ncu" ]; then
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_final_reference_arm.json 2>> gpurun_out/r2_bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras \
  --no-cpu-baseline --no-ref-gpu > gpurun_out/r2_launches_bench.log 2>&1
head -c 400 gpurun_out/r2_bench_final.json
fi
cap() {  # name, kernel regex, prof_run mode
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 \
    -o /tmp/$1 -f python tools/prof_run.py 33334 $3 > gpurun_out/$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_src.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
cap r2_pair2_energy k_pair_box2 inter
cap r2_pair2_force k_pair_box2 force
cap r2_nufft_spread k_nufft_spread recip
ls -la gpurun_out | tail -15
