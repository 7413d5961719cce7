#!/bin/bash
# Run on the GPU box via gpurun: FP64 peak, parity tests, quick timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./tools/fp64_peak | tee gpurun_out/fp64_peak.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python tools/quick_bench.py 10000 33334 2>&1 | tee gpurun_out/quick_bench.log | tail -20
