#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_zz_gpu_ffv1.py -x -q -m gpu 2>&1 | tail -3
for b in 32 64 128 256; do
  timeout 600 python benchmarks/ffv1_gpu_bench.py --frames $b --batch $b --reps 3 --context_model 2 --encode_only --grids auto 2>&1 | grep -v Warn | tail -1
done > gpurun_out/r02_ffv1_gpu_bench_model2.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_ffv1_gpu_bench_model2.jsonl'):
    try: d=json.loads(l); print('model 2 batch', d['batch'], round(d['device_frames_per_s']), d['bytes_per_frame'])
    except Exception as e: print(l[:300])
PY
