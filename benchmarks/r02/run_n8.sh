#!/bin/bash
# 8 GPUs of one box: host<->device ceiling with all GPUs copying at once, then the bench line at N=8 (end-to-end scaling)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
lscpu | head -30 > gpurun_out/r02_lscpu_n8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 benchmarks/pcie_probe.py > gpurun_out/r02_pcie_probe_n8.json 2> gpurun_out/r02_pcie_probe_n8.err
cat gpurun_out/r02_pcie_probe_n8.json | cut -c1-700
timeout 300 $TR --master-port 29532 benchmarks/pcie_probe.py --no-bind > gpurun_out/r02_pcie_probe_n8_nobind.json 2>> gpurun_out/r02_pcie_probe_n8.err
cat gpurun_out/r02_pcie_probe_n8_nobind.json | cut -c1-400
timeout 600 $TR --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
python -c "
import json;l=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1]);print(l['value'],l['e2e'])"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-paths --no-cpu > gpurun_out/r02_bench_n1_on_n8box.json 2> /dev/null
python -c "
import json;l=json.loads(open('gpurun_out/r02_bench_n1_on_n8box.json').read().strip().splitlines()[-1]);print(l['value'],l['e2e']['value'],l['e2e']['u8_mask']['value'])"
