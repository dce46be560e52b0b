#!/bin/bash
# 8 GPUs of one box: the bench line at N=8 after the round's changes (clock sampler through NVML, host pipeline)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_v3.json 2> gpurun_out/r02_bench_n8_v3.err
python -c "
import json;l=json.loads(open('gpurun_out/r02_bench_n8_v3.json').read().strip().splitlines()[-1]);print(l['value'],l['e2e']['value'],l['e2e']['u8_mask']['value'],l['clocks'],l['e2e'].get('host'))"
tail -2 gpurun_out/r02_bench_n8_v3.err
