#!/bin/bash
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8)
NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2
