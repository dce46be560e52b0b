#!/bin/bash
mkdir -p gpurun_out
timeout 300 python benchmarks/infill_breakdown.py > gpurun_out/r02_infill_breakdown.txt 2>&1; cat gpurun_out/r02_infill_breakdown.txt
timeout 600 python -m pytest tests/test_zz_gpu_ffv1.py -m gpu -q -x -k "chunk_reader or movie_steps or oracle_written" > gpurun_out/r02_run13_pytest.log 2>&1; tail -4 gpurun_out/r02_run13_pytest.log
