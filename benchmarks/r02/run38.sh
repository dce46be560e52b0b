#!/bin/bash
# round 2: host API, one stream per staging buffer (default) against upload / kernel / download streams (MDVT_HOST_PIPE=3)
for pipe in 2 3; do for slots in 2 3; do for chunk in 2 4; do echo -n "pipe=$pipe slots=$slots chunk=$chunk: "; MDVT_HOST_PIPE=$pipe MDVT_HOST_SLOTS=$slots MDVT_HOST_CHUNK=$chunk timeout 300 python bench.py --steps 10 --warmup 3 --no-paths --no-cpu 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['e2e']['value']), round(d['e2e']['u8_mask']['value']), d['e2e']['mask_checksum'], d['e2e']['u8_mask']['mask_checksum'])"; done; done; done > gpurun_out/r02_host_pipeline_sweep3.txt 2>&1
cat gpurun_out/r02_host_pipeline_sweep3.txt
