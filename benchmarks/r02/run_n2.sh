#!/bin/bash
# 2 GPUs of one box: the full GPU suite (the two torchrun 2-rank CLI tests run here), the randomised vrows test, bench at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_n2_v6.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_n2_v6.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_v6.json 2> gpurun_out/r02_bench_n2_v6.err
python -c "
import json;l=json.loads(open('gpurun_out/r02_bench_n2_v6.json').read().strip().splitlines()[-1]);print(l['value'],l['e2e']['value'],l['clocks'])"
tail -2 gpurun_out/r02_bench_n2_v6.err
