#!/bin/bash
# round 2: L2 policy of the z-buffer plane accesses in the generic loop: evict_last (built) / evict_normal / evict_first
for v in last normal first; do
  lib=$PWD/metric_depth_video_toolbox_b200/libmdvt_b200.so; [ $v != last ] && lib=$PWD/benchmarks/_variants/libmdvt_l2$v.so
  for sets in 1 2; do for dbg in 0 1; do echo "policy=$v MDVT_ZBUF_SETS=$sets MDVT_DEBUG=$dbg"; MDVT_B200_LIB=$lib MDVT_ZBUF_SETS=$sets MDVT_DEBUG=$dbg timeout 200 python benchmarks/quick_generic.py posed 2>&1 | tail -1 | cut -c60-130; done; done
done > gpurun_out/r02_generic_l2_policy.txt 2>&1
cat gpurun_out/r02_generic_l2_policy.txt
