#!/bin/bash
# r2v: what bounds the drain of the CTC head? timing probes (no stores / TMEM loads only); results of probe runs are wrong by design
set -u
mkdir -p gpurun_out
for p in 0 1 2; do
echo "probe $p"; CTCPS_HEAD_PROBE=$p timeout 120 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_head_gemm" -s 2 -c 2 --csv --log-file gpurun_out/r2v_probe$p.csv python tools/head_bench.py C2 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2v_probe$p.csv | cut -d, -f13- | tail -4
done
