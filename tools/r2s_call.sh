#!/bin/bash
# r2s: the CTC head with fp16 operands (3xFP16), bias staged in shared memory, transposed stores: tests, timing, one full capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ctc_head.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
timeout 200 python tools/head_bench.py C2 > gpurun_out/r2s_head_bench.log 2>&1; cat gpurun_out/r2s_head_bench.log
ncu --set full --clock-control none --import-source on -k regex:k_head_gemm -s 3 -c 1 -f -o gpurun_out/r2s_head_C2 python tools/head_bench.py C2 > gpurun_out/r2s_ncu_head.log 2>&1; echo "head capture rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_head|k_split|k_absmax" -c 12 --csv --log-file gpurun_out/r2s_head_launches.csv python tools/head_bench.py C2 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2s_head_launches.csv | cut -d, -f5,15- | head -14
