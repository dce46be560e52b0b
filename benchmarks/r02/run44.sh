#!/bin/bash
# round 2: lean resolve kernel of the frame loops: parity (whole GPU suite) + timing against the general kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for lean in 0 1; do echo "MDVT_RESOLVE_LEAN=$lean"; MDVT_RESOLVE_LEAN=$lean timeout 300 python benchmarks/quick_generic.py posed 2>&1 | tail -1; MDVT_RESOLVE_LEAN=$lean MDVT_DEBUG=1 timeout 300 python benchmarks/quick_generic.py posed 2>&1 | tail -1 | sed 's/^/   resolve only: /'; MDVT_RESOLVE_LEAN=$lean timeout 300 python benchmarks/quick_generic.py novel 2>&1 | tail -1; done > gpurun_out/r02_resolve_lean.txt 2>&1
cat gpurun_out/r02_resolve_lean.txt
