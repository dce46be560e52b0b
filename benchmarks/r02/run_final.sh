#!/bin/bash
# round 2, end-of-round verification on one B200: all GPU tests, smoke, both bench arms, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_final.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 600 gpurun_out/r02_bench_n1_final.json; tail -3 gpurun_out/r02_bench_n1_final.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1_final.json 2> gpurun_out/r02_bench_reference_n1_final.err; cut -c1-300 gpurun_out/r02_bench_reference_n1_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_under_ncu.log 2>&1
grep -c "mdvt" gpurun_out/r02_launches_bench_steps2.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stereo_rows -s 1 -c 1 -f -o gpurun_out/r02_stereo_rows_300f python benchmarks/quick_stereo.py 300 > gpurun_out/r02_rows_ncu.log 2>&1; tail -2 gpurun_out/r02_rows_ncu.log
timeout 300 python benchmarks/quick_generic.py both > gpurun_out/r02_paths_timings_final.txt 2>&1; cat gpurun_out/r02_paths_timings_final.txt
