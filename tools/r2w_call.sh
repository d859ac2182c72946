#!/bin/bash
# r2w: state after the windowing and the 3xFP16 head: GPU tests, full bench line, device timelines, drop-in probe
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2w_tests.log; cat gpurun_out/r2w_tests.log
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
timeout 200 python tools/dropin_probe.py C2 > gpurun_out/r2w_dropin_probe.json 2> gpurun_out/r2w_dropin_probe.err; cat gpurun_out/r2w_dropin_probe.json
for c in C2 C4; do
timeout 300 python tools/timeline.py --config $c --tag r2w_$c > gpurun_out/r2w_tl_$c.log 2>&1; echo "timeline $c rc=$?"
done
head -60 gpurun_out/r2w_C2_timeline.txt
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2w_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
for k in ("materialized_state", "pre_beam"):
    print(k, round(d[k]["value"]), round(d[k]["e2e"]["value"]))
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3), v["ctc_head_implementation"]) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict)})
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"]), "e2e", round(v["e2e"]["value"]), "score_ms", round(v["roofline"]["avg_launch_ms"], 4), "frac", round(v["roofline"]["frac"], 3), "mat", round(v["materialized_state"]["value"]), round(v["materialized_state"]["roofline"]["frac"], 3), "pre", round(v["pre_beam"]["value"]))
print("c5", {k: d["c5_job"].get(k) for k in ("value", "ms", "utterances_differing_from_aligned_transcript", "copies_of_an_utterance_agree", "hypotheses_checksum")})
P
