#!/bin/bash
# round 2: virtual-source-row kernel: 320 x 6 against 384 x 5 threads x columns
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" > gpurun_out/r02_vrows_tests_v7.log 2>&1; tail -3 gpurun_out/r02_vrows_tests_v7.log
timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v7_t320.txt 2>&1; cat gpurun_out/r02_vrows_timing_v7_t320.txt
MDVT_VROWS_T=384 timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v7_t384.txt 2>&1; cat gpurun_out/r02_vrows_timing_v7_t384.txt
MDVT_VROWS_T=384 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" 2>&1 | tail -2
