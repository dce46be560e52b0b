#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_dropin.py -q -x 2>&1 | tail -15
