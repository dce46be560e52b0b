#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_run4_breakdown.txt
for D in 0 1 2 4 6; do echo "MDVT_DEBUG=$D (1 no splat, 2 no resolve, 4 no re-arm)" >> gpurun_out/r02_run4_breakdown.txt; MDVT_DEBUG=$D timeout 300 python benchmarks/quick_generic.py posed >> gpurun_out/r02_run4_breakdown.txt 2>&1; done
cat gpurun_out/r02_run4_breakdown.txt
