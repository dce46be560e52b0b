#!/bin/bash
# round 2, verification on one B200 after the two-producer virtual-row kernel: all GPU tests, smoke, both bench arms,
# ncu full capture of the kernel (32 frames) and launch list of the bench command
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final5.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final5.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_final5.log 2>&1; tail -2 gpurun_out/r02_smoke_final5.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final5.json 2> gpurun_out/r02_bench_n1_final5.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1_final5.json 2> gpurun_out/r02_bench_reference_n1_final5.err; cut -c1-200 gpurun_out/r02_bench_reference_n1_final5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_steps2_v7.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_under_ncu_v5.log 2>&1
grep -c "mdvt" gpurun_out/r02_launches_bench_steps2_v7.csv
