#!/bin/bash
# r3b: padded frames of short utterances are not streamed: equivalence test, whole GPU suite, bench (C5 ragged job is where it shows)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_topk.py -q -m gpu -x -k "padded or skipping" 2>&1 | tail -6
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r3b_tests.log; cat gpurun_out/r3b_tests.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3b_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3), d["roofline"]["launches_timed"], d["roofline"]["launches_on_finished_batches_not_counted"],
      "| C1/C3/C4", [round(v["value"]) for v in d["configs"].values()], "| c5", round(d["c5_job"]["value"]), d["c5_job"]["hypotheses_checksum"], d["c5_job"]["utterances_differing_from_aligned_transcript"])
print("all rows scored:", round(d["finished_utterances_scored"]["value"]))
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
P
