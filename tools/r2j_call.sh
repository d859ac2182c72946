#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/head_bench.py C2 > gpurun_out/r2j_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_head_gemm -s 2 -c 1 -f -o gpurun_out/r2j_head python tools/head_bench.py C2 > gpurun_out/r2j_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2j_ncu.log
