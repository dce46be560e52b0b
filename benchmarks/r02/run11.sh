#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zz_gpu_ffv1.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02_run11_pytest.log 2>&1
tail -6 gpurun_out/r02_run11_pytest.log
for model in 0 1; do
  timeout 120 python benchmarks/ffv1_gpu_bench.py --frames 64 --batch 64 --reps 3 --grids auto --context_model $model --encode_only --decode > gpurun_out/r02_ffv1_bench_model${model}.jsonl 2> gpurun_out/r02_ffv1_bench_model${model}.err
done
cat gpurun_out/r02_ffv1_bench_model0.jsonl gpurun_out/r02_ffv1_bench_model1.jsonl
timeout 300 python benchmarks/movie_e2e.py 288 --green > gpurun_out/r02_movie_e2e_288_device.jsonl 2> gpurun_out/r02_movie_e2e_288_device.err; tail -1 gpurun_out/r02_movie_e2e_288_device.jsonl; tail -3 gpurun_out/r02_movie_e2e_288_device.err
timeout 300 python benchmarks/movie_e2e.py 288 --green --host-inputs > gpurun_out/r02_movie_e2e_288_hostin.jsonl 2> gpurun_out/r02_movie_e2e_288_hostin.err; tail -1 gpurun_out/r02_movie_e2e_288_hostin.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_splat|resolve_ckey|centroid" -s 3 -c 6 -o gpurun_out/r02_novel_v7 -f python benchmarks/novel_once.py > gpurun_out/r02_novel_ncu.log 2>&1
tail -2 gpurun_out/r02_novel_ncu.log
