#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_dropin_surface.py -m gpu -q > gpurun_out/r02_run3_pytest.log 2>&1
tail -40 gpurun_out/r02_run3_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r02_run3_smoke.log 2>&1; tail -3 gpurun_out/r02_run3_smoke.log
timeout 600 python benchmarks/quick_generic.py both > gpurun_out/r02_run3_timings.txt 2>&1
cat gpurun_out/r02_run3_timings.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stereo_conv -s 1 -c 1 -o gpurun_out/r02_conv_v5 -f python benchmarks/conv_once.py > gpurun_out/r02_conv_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_splat|resolve_ckey" -s 2 -c 4 -o gpurun_out/r02_generic_v9 -f python benchmarks/generic_once.py > gpurun_out/r02_generic_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_splat|resolve_ckey|centroid" -s 6 -c 6 -o gpurun_out/r02_novel_v6 -f python benchmarks/novel_once.py > gpurun_out/r02_novel_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
