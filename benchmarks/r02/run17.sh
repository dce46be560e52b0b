#!/bin/bash
# round 2: first run of the virtual-source-row convergence kernel: parity tests, then timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" > gpurun_out/r02_vrows_tests_v1.log 2>&1; tail -25 gpurun_out/r02_vrows_tests_v1.log
timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v1.txt 2>&1; cat gpurun_out/r02_vrows_timing_v1.txt
