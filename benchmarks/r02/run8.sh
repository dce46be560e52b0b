#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02_run8_pytest.log 2>&1
tail -4 gpurun_out/r02_run8_pytest.log
MDVT_ZBUF_SETS=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x -k "conv or cli or novel or project" > gpurun_out/r02_run8_pytest_sets2.log 2>&1
tail -3 gpurun_out/r02_run8_pytest_sets2.log
: > gpurun_out/r02_run8_timings.txt
for cfg in "1 100" "2 100" "2 60" "2 50" "2 40" "1 60"; do set -- $cfg; echo "MDVT_ZBUF_SETS=$1 MDVT_GRID_PCT=$2" >> gpurun_out/r02_run8_timings.txt; MDVT_ZBUF_SETS=$1 MDVT_GRID_PCT=$2 timeout 300 python benchmarks/quick_generic.py posed >> gpurun_out/r02_run8_timings.txt 2>&1; done
for D in 1 2; do echo "MDVT_DEBUG=$D" >> gpurun_out/r02_run8_timings.txt; MDVT_DEBUG=$D timeout 300 python benchmarks/quick_generic.py posed >> gpurun_out/r02_run8_timings.txt 2>&1; done
timeout 300 python benchmarks/quick_generic.py novel >> gpurun_out/r02_run8_timings.txt 2>&1
cat gpurun_out/r02_run8_timings.txt
