#!/bin/bash
# Everything about the device FFV1 codec that still needs a GPU measurement, in one gpurun call (about 5 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 700 -- 'bash benchmarks/ffv1_measure_all.sh'
# Results land in gpurun_out/ffv1_*.{log,jsonl,ncu-rep}.
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_zz_gpu_ffv1.py -q 2>&1 | tail -15 > gpurun_out/ffv1_tests.log
for model in 0 1; do
  timeout 60 python benchmarks/ffv1_gpu_bench.py --frames 64 --batch 64 --reps 3 --grids auto,32x16,16x16 --context_model $model --encode_only --decode \
    > gpurun_out/ffv1_bench_model${model}.jsonl 2> gpurun_out/ffv1_bench_model${model}.err
done
timeout 60 python benchmarks/ffv1_gpu_bench.py --frames 16 --batch 8 --reps 3 --grids auto > gpurun_out/ffv1_bench_batch8_files.jsonl 2>&1
MDVT_FFV1_WRITER=gpu timeout 120 python benchmarks/movie_e2e.py 288 --green > gpurun_out/ffv1_movie_e2e_288.jsonl 2> gpurun_out/ffv1_movie_e2e_288.err
MDVT_FFV1_WRITER=gpu timeout 120 python benchmarks/novel_e2e.py > gpurun_out/ffv1_novel_e2e_4k.jsonl 2> gpurun_out/ffv1_novel_e2e_4k.err
# the same two jobs with two packet-fed decoders per input (video_io.default_decoders): pays only with idle host cores
MDVT_READER_THREADS=2 MDVT_FFV1_WRITER=gpu timeout 120 python benchmarks/movie_e2e.py 288 --green > gpurun_out/ffv1_movie_e2e_288_dec2.jsonl 2>> gpurun_out/ffv1_movie_e2e_288.err
MDVT_READER_THREADS=2 MDVT_FFV1_WRITER=gpu timeout 120 python benchmarks/novel_e2e.py > gpurun_out/ffv1_novel_e2e_4k_dec2.jsonl 2>> gpurun_out/ffv1_novel_e2e_4k.err
timeout 90 ncu --set full --clock-control none --import-source on -k regex:ffv1_encode -c 1 -f -o gpurun_out/ffv1_encode_v2 \
  python benchmarks/ffv1_gpu_once.py > gpurun_out/ffv1_ncu_encode.log 2>&1
tail -n 3 gpurun_out/ffv1_tests.log
cat gpurun_out/ffv1_bench_model0.jsonl gpurun_out/ffv1_bench_model1.jsonl gpurun_out/ffv1_movie_e2e_288*.jsonl gpurun_out/ffv1_novel_e2e_4k*.jsonl
