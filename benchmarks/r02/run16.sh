#!/bin/bash
mkdir -p gpurun_out
MDVT_FFV1_ONCE_MODEL=1 MDVT_FFV1_ONCE_FRAMES=32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ffv1_encode|ffv1_decode" -c 2 -f -o gpurun_out/r02_ffv1_v3 python benchmarks/ffv1_gpu_once.py > gpurun_out/r02_ffv1_ncu.log 2>&1
tail -3 gpurun_out/r02_ffv1_ncu.log
timeout 200 python benchmarks/novel_e2e.py 48 > gpurun_out/r02_novel_e2e_4k_device.jsonl 2> gpurun_out/r02_novel_e2e_4k_device.err; tail -1 gpurun_out/r02_novel_e2e_4k_device.jsonl
