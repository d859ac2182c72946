#!/bin/bash
# r3p: per-tile scalars staged at the start of the tile (same commands as r3k / r3o: 0.2063 / 0.1947 ms at C2)
set -u
mkdir -p gpurun_out
for rep in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 --steps 20 --warmup 5 > gpurun_out/r3p_$rep.json 2> gpurun_out/r3p_$rep.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3p_$rep.json").read().strip().splitlines()[-1])
print("rep $rep:", "C2", round(d["value"]), "score ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
P
done
timeout 300 python bench.py --config C4 --no-cpu-baseline --single-mode --hidden-dim 0 --steps 10 --warmup 3 > gpurun_out/r3p_C4.json 2> gpurun_out/r3p_C4.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3p_C4.json").read().strip().splitlines()[-1])
print("C4:", round(d["value"]), "score ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
P
timeout 600 python -m pytest tests/test_gpu_fused_topk.py tests/test_gpu_beam_step.py tests/test_gpu_full_size.py -q -m gpu -x 2>&1 | tail -3
