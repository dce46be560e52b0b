#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zz_gpu_ffv1.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_run12_pytest.log 2>&1
tail -4 gpurun_out/r02_run12_pytest.log
MDVT_PROFILE_LOOP=1 timeout 300 python benchmarks/movie_e2e.py 288 --green > gpurun_out/r02_movie_e2e_288_device2.jsonl 2> gpurun_out/r02_movie_e2e_288_device2.err; tail -1 gpurun_out/r02_movie_e2e_288_device2.jsonl; grep "frame loop" gpurun_out/r02_movie_e2e_288_device2.err
