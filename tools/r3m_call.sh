#!/bin/bash
# r3m: where does the windowed scoring kernel spend its time?  full capture of a mid-decode launch at C2 and at C4
set -u
mkdir -p gpurun_out
for c in C2 C4; do
cmd="python bench.py --config $c --profile --steps 1 --warmup 1 --single-mode --hidden-dim 0 --no-cpu-baseline"
$cmd > gpurun_out/r3m_plain_$c.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_psi_full -s 75 -c 1 -f -o gpurun_out/r3m_psi_$c $cmd > gpurun_out/r3m_ncu_$c.log 2>&1
echo "$c capture rc=$?"
done
