#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zz_gpu_ffv1.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02_run14_pytest.log 2>&1
tail -4 gpurun_out/r02_run14_pytest.log
for C in 12 24; do
MDVT_PROFILE_LOOP=1 MDVT_E2E_CHUNK=$C timeout 300 python benchmarks/movie_e2e.py 288 --green > gpurun_out/r02_movie_e2e_288_device_c$C.jsonl 2> gpurun_out/r02_movie_e2e_288_device_c$C.err; tail -1 gpurun_out/r02_movie_e2e_288_device_c$C.jsonl; grep "frame loop" gpurun_out/r02_movie_e2e_288_device_c$C.err
done
MDVT_PROFILE_LOOP=1 timeout 300 python benchmarks/movie_e2e.py 960 --green > gpurun_out/r02_movie_e2e_960_device.jsonl 2> gpurun_out/r02_movie_e2e_960_device.err; tail -1 gpurun_out/r02_movie_e2e_960_device.jsonl; grep "frame loop" gpurun_out/r02_movie_e2e_960_device.err
timeout 300 python benchmarks/novel_e2e.py > gpurun_out/r02_novel_e2e_4k.jsonl 2> gpurun_out/r02_novel_e2e_4k.err; tail -1 gpurun_out/r02_novel_e2e_4k.jsonl
