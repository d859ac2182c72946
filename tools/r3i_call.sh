#!/bin/bash
# r3i: how far ahead of the GPU should the host run?  done_check_lag 1 / 2 / 3 / 4 at C2 (and C4), 20 steps each, alternating
set -u
mkdir -p gpurun_out
for rep in 1 2; do for lag in 1 2 3 4; do
timeout 300 python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 --steps 20 --warmup 5 --done-check-lag $lag > gpurun_out/r3i_lag${lag}_$rep.json 2> gpurun_out/r3i_lag${lag}_$rep.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3i_lag${lag}_$rep.json").read().strip().splitlines()[-1])
print("rep $rep lag $lag:", "C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "decode steps", d["config"]["decode_steps_per_utterance_batch"])
P
done; done
for lag in 1 3; do
timeout 300 python bench.py --config C4 --no-cpu-baseline --single-mode --hidden-dim 0 --steps 10 --warmup 3 --done-check-lag $lag > gpurun_out/r3i_C4_lag${lag}.json 2> gpurun_out/r3i_C4_lag${lag}.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3i_C4_lag${lag}.json").read().strip().splitlines()[-1])
print("C4 lag $lag:", round(d["value"]), "ms/step", round(d["ms_per_step"], 2))
P
done
