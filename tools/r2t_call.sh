#!/bin/bash
# r2t: GPU tests of the tree with the attention window (ctcps_score_window) and the 3xFP16 head
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/r2t_tests.log; cat gpurun_out/r2t_tests.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k window -s 2>&1 | grep -E "window_|passed|failed" > gpurun_out/r2t_window.log; cat gpurun_out/r2t_window.log
