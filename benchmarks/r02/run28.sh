#!/bin/bash
# round 2: compute-sanitizer over the virtual-row kernel (small frames), then the vrows tests incl. the randomised one
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do echo "== $tool"; timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=vrows python benchmarks/vrows_small.py 2>&1 | tail -12; done > gpurun_out/r02_vrows_sanitizer.txt 2>&1
cat gpurun_out/r02_vrows_sanitizer.txt | tail -45
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "vrows" 2>&1 | tail -3
