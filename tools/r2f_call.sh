#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_tests_full.log 2>&1
echo "full tests rc=$?"; tail -4 gpurun_out/r2f_tests_full.log
timeout 300 python tools/r2f_probe.py 2>&1 | tee gpurun_out/r2f_probe.log
timeout 600 python bench.py --no-cpu-baseline --c5-utterances 2048 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 --no-fuse-topk > gpurun_out/r2f_bench_unfused.json 2> gpurun_out/r2f_bench_unfused.err; echo "bench unfused rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
u = json.loads(open("gpurun_out/r2f_bench_unfused.json").read().strip().splitlines()[-1])
print("C2 fused", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", d["roofline"]["avg_launch_ms"], "frac", round(d["roofline"]["frac"], 3))
print("C2 unfused", round(u["value"]), "score_ms", u["roofline"]["avg_launch_ms"], "frac", round(u["roofline"]["frac"], 3))
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"]), "e2e", round(v["e2e"]["value"]), "score_ms", round(v["roofline"]["avg_launch_ms"], 4), "frac", round(v["roofline"]["frac"], 3), "mat", round(v["materialized_state"]["value"]), round(v["materialized_state"]["roofline"]["frac"], 3), "pre", round(v["pre_beam"]["value"]))
print("c5", {k: d["c5_job"][k] for k in ("value", "ms", "one_best_equals_transcripts", "hypotheses_checksum")})
P
