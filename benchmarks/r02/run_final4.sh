#!/bin/bash
# round 2, verification on one B200 after the two-producer virtual-row kernel: all GPU tests, smoke, both bench arms,
# ncu full capture of the kernel (32 frames) and launch list of the bench command
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final4.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final4.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_final4.log 2>&1; tail -2 gpurun_out/r02_smoke_final4.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:vrows -c 1 -o gpurun_out/r02_vrows_v14b_32f -f python benchmarks/quick_generic.py vrows > gpurun_out/r02_vrows_v14b_ncu.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final4.json 2> gpurun_out/r02_bench_n1_final4.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1_final4.json 2> gpurun_out/r02_bench_reference_n1_final4.err; cut -c1-200 gpurun_out/r02_bench_reference_n1_final4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_steps2_v6.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_under_ncu_v5.log 2>&1
grep -c "mdvt" gpurun_out/r02_launches_bench_steps2_v6.csv
