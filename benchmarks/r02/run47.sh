#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -3
