#!/bin/bash
cd /root/repo
for b in 64 128 256; do
  timeout 600 python benchmarks/ffv1_gpu_bench.py --frames $b --batch $b --reps 3 --context_model 1 --encode_only --grids auto 2>&1 | grep -v Warn | tail -2
done > gpurun_out/r02_ffv1_gpu_bench_batch_sweep.jsonl 2>&1
cat gpurun_out/r02_ffv1_gpu_bench_batch_sweep.jsonl | cut -c1-400
