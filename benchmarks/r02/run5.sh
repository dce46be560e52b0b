#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02_run5_pytest.log 2>&1
tail -4 gpurun_out/r02_run5_pytest.log
: > gpurun_out/r02_run5_breakdown.txt
for D in 0 1 2 4; do echo "MDVT_DEBUG=$D (1 no splat, 2 no resolve, 4 no re-arm)" >> gpurun_out/r02_run5_breakdown.txt; MDVT_DEBUG=$D timeout 300 python benchmarks/quick_generic.py posed >> gpurun_out/r02_run5_breakdown.txt 2>&1; done
timeout 300 python benchmarks/quick_generic.py both >> gpurun_out/r02_run5_breakdown.txt 2>&1
cat gpurun_out/r02_run5_breakdown.txt
