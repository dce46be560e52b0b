#!/bin/bash
# round 2: full GPU suite + smoke + bench after the convergence dispatch change (auto -> virtual-source-row kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_v4.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_v4.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_v4.log 2>&1; tail -2 gpurun_out/r02_smoke_v4.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_v4.json 2> gpurun_out/r02_bench_n1_v4.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_v4.json'))
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
for k,v in d['paths'].items(): print(k, round(v['ms_per_frame']*1e3,2), 'us', round(v['roofline']['frac'],3), round(v['e2e']['value']))
PY
tail -3 gpurun_out/r02_bench_n1_v4.err
