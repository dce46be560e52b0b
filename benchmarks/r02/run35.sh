#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_dropin.py tests/test_zz_gpu_ffv1.py -q -x 2>&1 | tail -4
