#!/bin/bash
# round 2: virtual-row kernel, 6 compute warps x 10 columns (MDVT_VROWS_T=192) against the default 5 x 12
MDVT_VROWS_T=192 timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v12_t192.txt 2>&1; cat gpurun_out/r02_vrows_timing_v12_t192.txt
MDVT_VROWS_T=192 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" 2>&1 | tail -2
