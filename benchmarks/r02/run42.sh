#!/bin/bash
python benchmarks/vrows_4k.py > gpurun_out/r02_vrows_4k.txt 2>&1; cat gpurun_out/r02_vrows_4k.txt
