#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ctc_head.py -q -m gpu -s 2>&1 | grep -E "max \||passed|failed|Error" | tail -12
timeout 120 python tools/head_bench.py C2 2>&1 | tee gpurun_out/r2i_head_bench.log
