#!/bin/bash
# round 2: compute-sanitizer memcheck over the two-producer virtual-row kernel: small frames incl. guard instantiations, then the vrows parity tests
cd /root/repo
{ echo "== memcheck, benchmarks/vrows_small.py"; timeout 600 compute-sanitizer --tool memcheck --kernel-name kns=vrows --print-limit 4 python benchmarks/vrows_small.py 2>&1 | tail -8
  echo "== memcheck, tests -k 'vrows or pans'"; timeout 1200 compute-sanitizer --tool memcheck --kernel-name kns=vrows --print-limit 4 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vrows or pans" 2>&1 | tail -8; } > gpurun_out/r02_vrows_sanitizer_v14.txt 2>&1
cat gpurun_out/r02_vrows_sanitizer_v14.txt
