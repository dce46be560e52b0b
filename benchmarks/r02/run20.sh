#!/bin/bash
# round 2: where the virtual-row kernel's time goes: phases switched off one at a time (results are wrong in these runs)
for s in 0 1 2 4 3 7; do echo "MDVT_VROWS_SKIP=$s"; MDVT_VROWS_SKIP=$s timeout 300 python benchmarks/quick_generic.py vrows 2>&1 | head -1; done > gpurun_out/r02_vrows_phase_skips_v7.txt 2>&1
cat gpurun_out/r02_vrows_phase_skips_v7.txt
