#!/bin/bash
# round 2, verification on one B200 after the virtual-row kernel: all GPU tests, smoke, both bench arms, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final2.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final2.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_final2.log 2>&1; tail -2 gpurun_out/r02_smoke_final2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_final2.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['roofline']['traffic_capture_matches_kernel_source'], d['e2e']['value'], d['clocks'])
for k,v in d['paths'].items():
    if 'ms_per_frame' in v: print(k, round(v['ms_per_frame']*1e3,2), 'us', round(v['roofline']['frac'],3), round(v['e2e']['value']))
    else: print(k, {kk: v[kk] for kk in v if kk in ('step4_frames_per_s','step5_frames_per_s','frames_per_s','error','frames')})
print(d.get('result_codec'))
PY
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1_final2.json 2> gpurun_out/r02_bench_reference_n1_final2.err; cut -c1-260 gpurun_out/r02_bench_reference_n1_final2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_steps2_v4.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_under_ncu_v4.log 2>&1
grep -c "mdvt" gpurun_out/r02_launches_bench_steps2_v4.csv
