#!/bin/bash
# r3d: where does k_beam_merge spend its time (C4: 54 us, C2: 23 us)?  one full capture each
set -u
mkdir -p gpurun_out
for c in C4 C2; do
cmd="python bench.py --config $c --profile --steps 1 --warmup 1 --single-mode --hidden-dim 0 --no-cpu-baseline"
$cmd > gpurun_out/r3d_plain_$c.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_beam_merge -s 40 -c 1 -f -o gpurun_out/r3d_merge_$c $cmd > gpurun_out/r3d_ncu_$c.log 2>&1
echo "$c capture rc=$?"
done
