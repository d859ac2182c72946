#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ctc_head.py -q -m gpu -x -s > gpurun_out/r2g_head_tests.log 2>&1
echo "head tests rc=$?"; grep -E "max \||passed|failed|Error|error" gpurun_out/r2g_head_tests.log | tail -20
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
