#!/bin/bash
# round 2: compute-sanitizer memcheck over the GPU parity tests (all kernels of the library)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r02_memcheck_gpu_parity.log 2>&1
grep -v "Host Frame" gpurun_out/r02_memcheck_gpu_parity.log | tail -25
