#!/bin/bash
# round 2: DRAM traffic of the generic two-lane loop in its steady state (no cache flush between kernels)
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:"project_splat|resolve_ckey" -s 32 -c 32 --csv --log-file gpurun_out/r02_generic_steady_dram.csv python benchmarks/generic_steady.py > gpurun_out/r02_generic_steady.log 2>&1
tail -2 gpurun_out/r02_generic_steady.log; head -c 3000 gpurun_out/r02_generic_steady_dram.csv | tail -c 1500
