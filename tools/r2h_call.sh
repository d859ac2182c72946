#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_ctc_head.py -q -m gpu -s 2>&1 | grep -E "max \||passed|failed" | tail -12
timeout 120 python tools/head_bench.py C2 2>&1 | tee gpurun_out/r2h_head_bench.log
timeout 120 python tools/head_bench.py C2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_head|k_split|k_init|gemm|cutlass|sm100|nvjet" -c 40 --csv --log-file gpurun_out/r2h_head_launches.csv python tools/head_bench.py C2 > gpurun_out/r2h_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2h_head_launches.csv | head -12
