#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2m_tests_full.log 2>&1
echo "full tests rc=$?"; tail -4 gpurun_out/r2m_tests_full.log
( time timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
CTCPS_HEAD=cublas timeout 300 python bench.py --no-cpu-baseline --extra-configs "" --c5-utterances 0 --no-drop-in --pre-beam 0 > gpurun_out/r2m_bench_cublas_head.json 2> gpurun_out/r2m_bench_cublas_head.err; echo "bench cublas-head rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
c = json.loads(open("gpurun_out/r2m_bench_cublas_head.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
for k in ("materialized_state", "pre_beam"):
    print(k, round(d[k]["value"]), round(d[k]["e2e"]["value"]))
print("hidden tcgen05", {k: (round(v["value"]), round(v["ctc_head_ms"], 3), v["ctc_head_implementation"]) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict)})
print("hidden cublas ", {k: (round(v["value"]), round(v["ctc_head_ms"], 3), v["ctc_head_implementation"]) for k, v in c["e2e_from_hidden"].items() if isinstance(v, dict)})
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"]), "e2e", round(v["e2e"]["value"]), "score_ms", round(v["roofline"]["avg_launch_ms"], 4), "frac", round(v["roofline"]["frac"], 3), "mat", round(v["materialized_state"]["value"]), round(v["materialized_state"]["roofline"]["frac"], 3), "pre", round(v["pre_beam"]["value"]))
print("c5", {k: d["c5_job"][k] for k in ("value", "ms", "one_best_equals_transcripts", "hypotheses_checksum")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"][:60])
P
