#!/bin/bash
# round 2: virtual-row kernel, columns per batch (shared-memory loads in flight per thread): 4 / 6 / 12
for nb in 4 6 12; do echo "NB=$nb"; MDVT_B200_LIB=$PWD/benchmarks/_variants/libmdvt_nb$nb.so timeout 300 python benchmarks/quick_generic.py vrows 2>&1; done > gpurun_out/r02_vrows_nb_sweep.txt 2>&1
cat gpurun_out/r02_vrows_nb_sweep.txt
