#!/bin/bash
# round 2, GPU call 1: ray-form arithmetic in the generic splat and the convergence kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_run1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "project_splat or conv or full_size or unproject or stereo_rows_equals" > gpurun_out/r02_run1_pytest_a.log 2>&1
tail -5 gpurun_out/r02_run1_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "novel or render or cli_stereo_rerender_conv or movie" > gpurun_out/r02_run1_pytest_b.log 2>&1
tail -5 gpurun_out/r02_run1_pytest_b.log
: > gpurun_out/r02_run1_timings.txt
for U in 1 2 4; do echo "MDVT_CONV_U=$U" >> gpurun_out/r02_run1_timings.txt; MDVT_CONV_U=$U timeout 300 python benchmarks/quick_generic.py stereo >> gpurun_out/r02_run1_timings.txt 2>&1; done
timeout 300 python benchmarks/quick_generic.py novel >> gpurun_out/r02_run1_timings.txt 2>&1
cat gpurun_out/r02_run1_timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stereo_conv -s 2 -c 1 -o gpurun_out/r02_conv_v4 -f python benchmarks/conv_once.py > gpurun_out/r02_conv_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_splat|resolve_rows" -s 2 -c 4 -o gpurun_out/r02_generic_v8 -f python benchmarks/generic_once.py > gpurun_out/r02_generic_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
