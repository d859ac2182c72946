#!/bin/bash
# ncu launch list of one decode + one full capture of the lazy scoring kernel, per config.
#   gpurun --timeout 1200 -- 'bash tools/prof_psi.sh r2c C2 C4'
# Output: gpurun_out/<tag>_launches_<cfg>.csv, gpurun_out/<tag>_psi_<cfg>.ncu-rep (+ .raw.csv).  Nothing printed here is a bench value.
set -u
tag=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  extra=""
  case $cfg in *-unfused) extra="--no-fuse-topk"; c=${cfg%-unfused};; *) c=$cfg;; esac
  cmd="python bench.py --config $c --profile --steps 1 --warmup 1 --single-mode --hidden-dim 0 --no-cpu-baseline $extra"
  $cmd > gpurun_out/${tag}_plain_$cfg.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_$cfg.csv $cmd > gpurun_out/${tag}_ncu1_$cfg.log 2>&1
  echo "$cfg launch list rc=$?"
  $cmd > gpurun_out/${tag}_plain2_$cfg.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_psi_full -s 60 -c 1 -f -o gpurun_out/${tag}_psi_$cfg $cmd > gpurun_out/${tag}_ncu2_$cfg.log 2>&1
  echo "$cfg full capture rc=$?"
done
for f in gpurun_out/${tag}_launches_*.csv; do echo "== $f"; python tools/launch_summary.py $f 2>/dev/null | head -16; done
