#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_zz_gpu_ffv1.py -x -q -m gpu 2>&1 | tail -3
for m in 1 2; do for b in 64 128; do
  timeout 600 python benchmarks/ffv1_gpu_bench.py --frames $b --batch $b --reps 3 --context_model $m --encode_only --decode --grids auto 2>&1 | grep -v Warn | tail -2
done; done > gpurun_out/r02_ffv1_gpu_bench_model2_decode.jsonl 2>&1
cut -c1-330 gpurun_out/r02_ffv1_gpu_bench_model2_decode.jsonl
