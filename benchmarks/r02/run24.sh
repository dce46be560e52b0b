#!/bin/bash
# round 2: generic loop, plane sets (1 / 2) x kernels switched off (MDVT_DEBUG 1 = no splat, 2 = no resolve): is the resolve's plane read an L2 hit?
for sets in 1 2; do for dbg in 0 1 2; do echo "MDVT_ZBUF_SETS=$sets MDVT_DEBUG=$dbg"; MDVT_ZBUF_SETS=$sets MDVT_DEBUG=$dbg timeout 200 python benchmarks/quick_generic.py posed 2>&1 | tail -1; done; done > gpurun_out/r02_generic_sets_debug_matrix.txt 2>&1
cat gpurun_out/r02_generic_sets_debug_matrix.txt
