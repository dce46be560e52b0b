#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02_run9_pytest.log 2>&1
tail -4 gpurun_out/r02_run9_pytest.log
: > gpurun_out/r02_run9_timings.txt
for S in 2 1; do echo "MDVT_ZBUF_SETS=$S" >> gpurun_out/r02_run9_timings.txt; MDVT_ZBUF_SETS=$S timeout 300 python benchmarks/quick_generic.py both >> gpurun_out/r02_run9_timings.txt 2>&1; done
cat gpurun_out/r02_run9_timings.txt
