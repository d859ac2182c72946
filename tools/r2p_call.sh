#!/bin/bash
# r2p: A/B of the materialised kernel's hypothesis-group width on the small shape (C1), then the full bench line again
set -u
mkdir -p gpurun_out
for g in 5 3 2; do
CTCPS_SCORE_MAX_GROUP=$g timeout 200 python bench.py --config C1 --state materialized --single-mode --no-cpu-baseline --hidden-dim 0 --steps 3 > gpurun_out/r2p_C1_mat_g$g.json 2> gpurun_out/r2p_C1_mat_g$g.err
python - <<P
import json
d = json.loads(open("gpurun_out/r2p_C1_mat_g$g.json").read().strip().splitlines()[-1])
print("C1 materialised, max group $g:", round(d["value"]), "utt/s, score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
P
done
for g in 5 2; do
CTCPS_SCORE_MAX_GROUP=$g timeout 300 python bench.py --config C2 --state materialized --single-mode --no-cpu-baseline --hidden-dim 0 --steps 2 --warmup 1 > gpurun_out/r2p_C2_mat_g$g.json 2> gpurun_out/r2p_C2_mat_g$g.err
python - <<P
import json
d = json.loads(open("gpurun_out/r2p_C2_mat_g$g.json").read().strip().splitlines()[-1])
print("C2 materialised, max group $g:", round(d["value"]), "utt/s, score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
P
done
( time timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3), v["ctc_head_implementation"]) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict)})
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
print("c5", {k: d["c5_job"][k] for k in ("value", "ms", "utterances_differing_from_aligned_transcript", "copies_of_an_utterance_agree", "hypotheses_checksum")})
P
