#!/bin/bash
# r3k: L2 evict-first priority for the posteriors stream of the lazy scoring kernel: A/B at C2 and C4 (alternating), then tests
set -u
mkdir -p gpurun_out
for rep in 1 2; do for ef in 0 1; do
CTCPS_PSI_EVICT_FIRST=$ef timeout 300 python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 --steps 20 --warmup 5 > gpurun_out/r3k_ef${ef}_$rep.json 2> gpurun_out/r3k_ef${ef}_$rep.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3k_ef${ef}_$rep.json").read().strip().splitlines()[-1])
print("rep $rep evict_first=$ef:", "C2", round(d["value"]), "score ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
P
done; done
for ef in 0 1; do
CTCPS_PSI_EVICT_FIRST=$ef timeout 300 python bench.py --config C4 --no-cpu-baseline --single-mode --hidden-dim 0 --steps 10 --warmup 3 > gpurun_out/r3k_C4_ef${ef}.json 2> gpurun_out/r3k_C4_ef${ef}.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3k_C4_ef${ef}.json").read().strip().splitlines()[-1])
print("C4 evict_first=$ef:", round(d["value"]), "score ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
P
done
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
