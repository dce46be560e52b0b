#!/bin/bash
# round 2: generic loop with / without an L2 persisting window on the z-buffer planes
for p in 0 1; do for dbg in 0 1 2; do echo "MDVT_L2_PERSIST=$p MDVT_DEBUG=$dbg"; MDVT_L2_PERSIST=$p MDVT_DEBUG=$dbg timeout 200 python benchmarks/quick_generic.py posed 2>&1 | tail -1; done; done > gpurun_out/r02_generic_l2_persist.txt 2>&1
echo "sets=1 persist=1"; MDVT_ZBUF_SETS=1 MDVT_L2_PERSIST=1 timeout 200 python benchmarks/quick_generic.py posed 2>&1 | tail -1 >> gpurun_out/r02_generic_l2_persist.txt
cat gpurun_out/r02_generic_l2_persist.txt
