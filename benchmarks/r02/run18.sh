#!/bin/bash
# round 2: virtual-source-row kernel v11 (float32 staircase, ballot runs): parity, timing, ncu capture with source
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" > gpurun_out/r02_vrows_tests_v11.log 2>&1; tail -5 gpurun_out/r02_vrows_tests_v11.log
timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v11.txt 2>&1; cat gpurun_out/r02_vrows_timing_v11.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vrows -s 1 -c 1 -f -o gpurun_out/r02_vrows_v11 python benchmarks/conv_once.py vrows 5.0 > gpurun_out/r02_vrows_ncu_v11.log 2>&1; tail -2 gpurun_out/r02_vrows_ncu_v11.log
