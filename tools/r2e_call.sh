#!/bin/bash
# r2e: new tests (N3, pre-beam without KV cache), pscan fp64 numbers, the full new bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2e_tests_full.log 2>&1
echo "full tests rc=$?"; tail -4 gpurun_out/r2e_tests_full.log
timeout 300 python -m pytest tests/test_gpu_select_pscan.py tests/test_gpu_full_size.py -q -m gpu -s -k "sequential_scan or full_size_properties" 2>&1 | grep -E "pscan-vs-fp64|survivors vs fp64|passed|failed" > gpurun_out/r2e_pscan_fp64.log
cat gpurun_out/r2e_pscan_fp64.log
( time timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -3 gpurun_out/r2e_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
for k in ("materialized_state", "pre_beam"):
    print(k, round(d[k]["value"]), round(d[k]["e2e"]["value"]))
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3)) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict)})
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"]), "e2e", round(v["e2e"]["value"]), "frac", round(v["roofline"]["frac"], 3), "mat", round(v["materialized_state"]["value"]), round(v["materialized_state"]["roofline"]["frac"], 3), "pre", round(v["pre_beam"]["value"]))
print("c5", {k: d["c5_job"][k] for k in ("value", "ms", "wall_s", "one_best_equals_transcripts", "hypotheses_checksum", "batches_this_rank")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
P
