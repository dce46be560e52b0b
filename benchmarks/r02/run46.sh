#!/bin/bash
# round 2: generic loop, lanes split by frame (default) against lanes split by view (MDVT_LANES=views: half the live planes)
for lanes in frames views; do echo "MDVT_LANES=$lanes"; MDVT_LANES=$lanes timeout 300 python benchmarks/quick_generic.py posed 2>&1 | tail -1; done > gpurun_out/r02_generic_lanes_by_view.txt 2>&1
MDVT_LANES=views timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -2 >> gpurun_out/r02_generic_lanes_by_view.txt
cat gpurun_out/r02_generic_lanes_by_view.txt
