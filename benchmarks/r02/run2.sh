#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_run2_timings.txt
for T in 128 256 384 512; do echo "MDVT_CONV_THREADS=$T" >> gpurun_out/r02_run2_timings.txt; MDVT_CONV_THREADS=$T timeout 300 python benchmarks/quick_generic.py conv >> gpurun_out/r02_run2_timings.txt 2>&1; done
cat gpurun_out/r02_run2_timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stereo_conv -s 1 -c 1 -o gpurun_out/r02_conv_v4 -f python benchmarks/conv_once.py > gpurun_out/r02_conv_ncu.log 2>&1
tail -3 gpurun_out/r02_conv_ncu.log
