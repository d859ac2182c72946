#!/bin/bash
# r3f: the lazy scoring kernel streams only the chunks whose lin weights are nonzero: equivalence tests, whole suite, A/B
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_topk.py -q -m gpu -x -k "nonzero_weight" -s 2>&1 | grep -E "chunks|passed|failed|Error|assert" | tail -12
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r3f_tests.log; cat gpurun_out/r3f_tests.log
for on in 0 1; do
CTCPS_FRAME_WINDOW=$on timeout 400 python bench.py --no-cpu-baseline --no-drop-in --pre-beam 0 --hidden-dim 0 --steps 10 --warmup 3 > gpurun_out/r3f_bench_win$on.json 2> gpurun_out/r3f_bench_win$on.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3f_bench_win$on.json").read().strip().splitlines()[-1])
print("frame_window=$on:", "C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3),
      "| C1/C3/C4", [round(v["value"]) for v in d["configs"].values()], [round(v["roofline"]["avg_launch_ms"], 4) for v in d["configs"].values()], "| c5", round(d["c5_job"]["value"]), d["c5_job"]["hypotheses_checksum"], d["c5_job"]["utterances_differing_from_aligned_transcript"])
P
done
