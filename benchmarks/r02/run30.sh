#!/bin/bash
# round 2: synccheck / memcheck of the virtual-row kernel after the warp reconvergence fix, vrows tests, timing
for tool in synccheck memcheck; do echo "== $tool"; timeout 600 compute-sanitizer --tool $tool --kernel-name kns=vrows --print-limit 4 python benchmarks/vrows_small.py 2>&1 | tail -8; done > gpurun_out/r02_vrows_sanitizer_v13.txt 2>&1
cat gpurun_out/r02_vrows_sanitizer_v13.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" 2>&1 | tail -2
timeout 300 python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_timing_v13.txt 2>&1; cat gpurun_out/r02_vrows_timing_v13.txt
