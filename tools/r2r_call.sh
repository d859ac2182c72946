#!/bin/bash
# r2r: state of the tree after the container was re-created: GPU tests, the full bench line, the reference arm, launch list + full captures at C2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2r_tests_full.log 2>&1
echo "full tests rc=$?"; tail -4 gpurun_out/r2r_tests_full.log
( time timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 400 python bench.py --impl reference > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err ) 2>&1 | grep real; echo "reference rc=$?"
timeout 200 python tools/head_bench.py C2 > gpurun_out/r2r_head_bench.log 2>&1; cat gpurun_out/r2r_head_bench.log
bash tools/prof_psi.sh r2r C2
# one full capture of the CTC head
cmd="python tools/head_bench.py C2"
ncu --set full --clock-control none --import-source on -k regex:k_head_gemm -s 3 -c 1 -f -o gpurun_out/r2r_head_C2 $cmd > gpurun_out/r2r_ncu_head.log 2>&1; echo "head capture rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
for k in ("materialized_state", "pre_beam"):
    print(k, round(d[k]["value"]), round(d[k]["e2e"]["value"]))
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3), v["ctc_head_implementation"]) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict)})
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"]), "e2e", round(v["e2e"]["value"]), "score_ms", round(v["roofline"]["avg_launch_ms"], 4), "frac", round(v["roofline"]["frac"], 3), "mat", round(v["materialized_state"]["value"]), round(v["materialized_state"]["roofline"]["frac"], 3), "pre", round(v["pre_beam"]["value"]))
print("c5", {k: d["c5_job"].get(k) for k in ("value", "ms", "utterances_differing_from_aligned_transcript", "copies_of_an_utterance_agree", "hypotheses_checksum")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"][:60])
P
