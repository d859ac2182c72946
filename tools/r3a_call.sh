#!/bin/bash
# r3a: finished utterances are not scored by the native loop: equivalence tests, then A/B of the bench step and the ragged C5 job
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_topk.py tests/test_gpu_beam_step.py tests/test_gpu_full_size.py -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r3a_tests.log; cat gpurun_out/r3a_tests.log
for on in 0 1; do
CTCPS_SKIP_DONE=$on timeout 400 python bench.py --no-cpu-baseline --no-drop-in --pre-beam 0 --hidden-dim 0 --steps 10 --warmup 3 > gpurun_out/r3a_bench_skip$on.json 2> gpurun_out/r3a_bench_skip$on.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3a_bench_skip$on.json").read().strip().splitlines()[-1])
print("skip_done=$on:", "C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3),
      "| C1/C3/C4", [round(v["value"]) for v in d["configs"].values()], "| c5", round(d["c5_job"]["value"]), d["c5_job"]["hypotheses_checksum"], d["c5_job"]["utterances_differing_from_aligned_transcript"])
P
done
