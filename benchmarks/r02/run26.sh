#!/bin/bash
# round 2: virtual-row kernel v12 (cull-free pass 1 where the near plane allows it, phase B tidied): parity + timing
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" > gpurun_out/r02_vrows_tests_v12.log 2>&1; tail -3 gpurun_out/r02_vrows_tests_v12.log
timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v12.txt 2>&1; cat gpurun_out/r02_vrows_timing_v12.txt
