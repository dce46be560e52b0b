#!/bin/bash
# round 2: compute-sanitizer memcheck over the drop-in and FFV1 GPU tests
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dropin.py tests/test_zz_gpu_ffv1.py -q -x --deselect tests/test_zz_gpu_ffv1.py::test_cli_stereo_rerender_gpu_writer_torchrun_two_ranks > gpurun_out/r02_memcheck_gpu_dropin_ffv1.log 2>&1
grep -v "Host Frame" gpurun_out/r02_memcheck_gpu_dropin_ffv1.log | grep -v "^tests/\|RuntimeWarning\|near_half" | tail -25
