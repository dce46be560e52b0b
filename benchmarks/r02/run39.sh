#!/bin/bash
# round 2: virtual-row kernel, 4 compute warps x 15 columns (MDVT_VROWS_T=128) against the default 5 x 12
for t in 0 128; do echo "MDVT_VROWS_T=$t"; MDVT_VROWS_T=$t timeout 300 python benchmarks/quick_generic.py vrows 2>&1; done > gpurun_out/r02_vrows_timing_v13_t128.txt; cat gpurun_out/r02_vrows_timing_v13_t128.txt
MDVT_VROWS_T=128 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" 2>&1 | tail -2
