#!/bin/bash
# round 2: does a fifth resident CTA per SM help the virtual-row kernel? 1280x720 (39 KB of shared memory per CTA), default build (<= 80 registers: 4 CTAs) against a build capped at 64 registers (5 CTAs)
echo "default build"; python benchmarks/vrows_720p.py
echo "64-register build"; MDVT_B200_LIB=$PWD/benchmarks/_variants/libmdvt_vrows_r64.so python benchmarks/vrows_720p.py
