#!/bin/bash
# round 2: host API pipeline depth: staging slots x frames per slot (second sweep: small chunks)
for slots in 2 3; do for chunk in 1 2 3 4 6; do echo -n "slots=$slots chunk=$chunk: "; MDVT_HOST_SLOTS=$slots MDVT_HOST_CHUNK=$chunk timeout 300 python bench.py --steps 10 --warmup 3 --no-paths --no-cpu 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['e2e']['value']), round(d['e2e']['u8_mask']['value']))"; done; done > gpurun_out/r02_host_pipeline_sweep2.txt 2>&1
cat gpurun_out/r02_host_pipeline_sweep2.txt
