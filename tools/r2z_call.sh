#!/bin/bash
# r2z: does binding the single rank to the GPU's NUMA node change the e2e (H2D-bound) number?  + device timelines of C1 / C3
set -u
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
for f in "" "--no-numa-bind"; do
timeout 300 python bench.py --no-cpu-baseline --extra-configs "" --c5-utterances 0 --no-drop-in --pre-beam 0 --hidden-dim 0 --steps 10 --warmup 3 $f > gpurun_out/r2z_bench$f.json 2> gpurun_out/r2z_bench$f.err
python - <<P
import json
d = json.loads(open("gpurun_out/r2z_bench$f.json").read().strip().splitlines()[-1])
print("numa flag '$f':", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "bound cpus", d["config"].get("numa_bound_cpus"))
P
done
for c in C1 C3; do
timeout 300 python tools/timeline.py --config $c --tag r2z_$c > gpurun_out/r2z_tl_$c.log 2>&1; echo "timeline $c rc=$?"; head -16 gpurun_out/r2z_${c}_timeline.txt
done
