#!/bin/bash
# r2u: the CTC head with the normalisation fused into the GEMM kernel (bands + grid barrier, rows read back from L2)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ctc_head.py -q -m gpu -x -s 2>&1 | grep -E "error|Error|passed|failed|max \|" | tail -15 > gpurun_out/r2u_tests.log; cat gpurun_out/r2u_tests.log
timeout 120 python tools/head_bench.py C2 > gpurun_out/r2u_head_bench.log 2>&1; cat gpurun_out/r2u_head_bench.log
timeout 120 python tools/head_bench.py C4 2>&1 | head -1; timeout 120 python tools/head_bench.py C1 2>&1 | head -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_head_gemm -s 3 -c 1 -f -o gpurun_out/r2u_head_C2 python tools/head_bench.py C2 > gpurun_out/r2u_ncu_head.log 2>&1; echo "head capture rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_head|k_split|k_absmax" -c 8 --csv --log-file gpurun_out/r2u_head_launches.csv python tools/head_bench.py C2 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2u_head_launches.csv | cut -d, -f5,13- | head -24
