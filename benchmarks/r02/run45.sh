#!/bin/bash
# round 2: generic loop, cost per pixel against the size of the planes (all kernels / resolve only / splat only)
for dbg in 0 1 2; do echo "MDVT_DEBUG=$dbg"; MDVT_DEBUG=$dbg timeout 300 python benchmarks/generic_sizes.py; done > gpurun_out/r02_generic_sizes.txt 2>&1; cat gpurun_out/r02_generic_sizes.txt
