#!/bin/bash
# the bench leg of run_final2.sh (the box has no /usr/bin/time: wall clock through date)
mkdir -p gpurun_out
t0=$(date +%s.%N)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err
t1=$(date +%s.%N); echo "bench.py wall seconds: $(echo "$t1 - $t0" | bc)"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final2.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['roofline']['traffic_capture_matches_kernel_source'], d['e2e']['value'], d['clocks'])
for k,v in d['paths'].items():
    if 'ms_per_frame' in v: print(k, round(v['ms_per_frame']*1e3,2), 'us', round(v['roofline']['frac'],3), round(v['e2e']['value']))
    else: print(k, {kk: v[kk] for kk in v if kk in ('step4_frames_per_s','step5_frames_per_s','frames_per_s','error','frames')})
print(d.get('result_codec'))
print(d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r02_bench_n1_final2.err
