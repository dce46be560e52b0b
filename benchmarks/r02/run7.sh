#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_surface.py -m gpu -q -x -k "pack_mask or encode or rows_bit_exact or dropin or normals or mesh" > gpurun_out/r02_run7_pytest.log 2>&1
tail -4 gpurun_out/r02_run7_pytest.log
timeout 300 python benchmarks/pcie_probe.py > gpurun_out/r02_pcie_probe_n1.json 2>gpurun_out/r02_pcie_probe_n1.err; cat gpurun_out/r02_pcie_probe_n1.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-paths > gpurun_out/r02_bench_n1_v2.json 2> gpurun_out/r02_bench_n1_v2.err
python -c "
import json;l=json.loads(open('gpurun_out/r02_bench_n1_v2.json').read().strip().splitlines()[-1]);print(l['value'],l['e2e'])"
tail -3 gpurun_out/r02_bench_n1_v2.err
