#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_splat|resolve_ckey|centroid" -s 9 -c 6 -o gpurun_out/r02_novel_v7 -f python benchmarks/novel_once.py > gpurun_out/r02_novel_ncu.log 2>&1
tail -2 gpurun_out/r02_novel_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_splat|resolve_ckey" -s 4 -c 4 -o gpurun_out/r02_generic_v10 -f python benchmarks/generic_once.py > gpurun_out/r02_generic_ncu.log 2>&1
tail -2 gpurun_out/r02_generic_ncu.log
for P in 100 70 50; do echo "MDVT_GRID_PCT=$P"; MDVT_GRID_PCT=$P timeout 300 python benchmarks/quick_generic.py novel; done > gpurun_out/r02_run10_timings.txt 2>&1
cat gpurun_out/r02_run10_timings.txt
