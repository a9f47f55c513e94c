#!/bin/bash
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "${1:-golden or stagewise or cfg3}" 2>&1 | tail -4)
timeout 300 python tools/gpu_timing.py 2>&1 | tail -5
