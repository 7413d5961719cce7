#!/bin/bash
# Full ncu captures of the dominant kernels, exported to CSV on the GPU box (the .ncu-rep
# files are too large to travel back through gpurun_out/).
mkdir -p gpurun_out
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
