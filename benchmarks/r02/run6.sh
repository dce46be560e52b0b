#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_surface.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02_run6_pytest.log 2>&1
tail -4 gpurun_out/r02_run6_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_v1.json 2> gpurun_out/r02_bench_n1_v1.err
tail -c 3000 gpurun_out/r02_bench_n1_v1.json; tail -5 gpurun_out/r02_bench_n1_v1.err
