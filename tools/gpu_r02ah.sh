#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | tail -30
