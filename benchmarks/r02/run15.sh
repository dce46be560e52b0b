#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_ffv1.py -m gpu -q -x -k "chunk_reader or movie_steps or cli" > gpurun_out/r02_run15_pytest.log 2>&1
tail -3 gpurun_out/r02_run15_pytest.log
for C in 12 24; do
MDVT_PROFILE_LOOP=1 MDVT_E2E_CHUNK=$C timeout 300 python benchmarks/movie_e2e.py 960 --green > gpurun_out/r02_movie_e2e_960_device_c$C.jsonl 2> gpurun_out/r02_movie_e2e_960_device_c$C.err; tail -1 gpurun_out/r02_movie_e2e_960_device_c$C.jsonl; grep "frame loop" gpurun_out/r02_movie_e2e_960_device_c$C.err
done
