#!/bin/bash
timeout 600 compute-sanitizer --tool synccheck --kernel-name kns=vrows --print-limit 3 python benchmarks/vrows_small.py > gpurun_out/r02_vrows_synccheck.txt 2>&1
grep -v "Host Frame" gpurun_out/r02_vrows_synccheck.txt | head -30
