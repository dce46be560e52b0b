#!/bin/bash
python benchmarks/l2_capacity_probe.py > gpurun_out/r02_l2_capacity_probe.txt 2>&1; cat gpurun_out/r02_l2_capacity_probe.txt
