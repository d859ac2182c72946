#!/bin/bash
# r3q: final tree: whole GPU suite, smoke, the bench line with the driver's flags
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r3q_tests.log; cat gpurun_out/r3q_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3q_bench.json 2> gpurun_out/r3q_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<P
import json
d = json.loads(open("gpurun_out/r3q_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "roofline achieved", round(r["achieved"]), "frac", round(r["frac"], 3), "avg ms", round(r["avg_launch_ms"], 4), "streamed frac", round(r.get("streamed_fraction_of_all_frames", 0), 3), "launches", d["gpu_launches"])
e = d["every_row_every_frame"]
print("every row every frame:", round(e["value"]), "frac", round(e["roofline"]["frac"], 3), "ms", round(e["roofline"]["avg_launch_ms"], 4))
print("mat", round(d["materialized_state"]["value"]), round(d["materialized_state"]["roofline"]["frac"], 3), "pre", round(d["pre_beam"]["value"]))
print("configs", {k: (round(v["value"]), round(v["e2e"]["value"]), round(v["roofline"]["frac"], 3), round(v["materialized_state"]["roofline"]["frac"], 3), round(v["pre_beam"]["value"])) for k, v in d["configs"].items()})
h = d["e2e_from_hidden"]
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3)) for k, v in h.items() if isinstance(v, dict) and "value" in v}, "head roofline", round(h["head_roofline"]["frac"], 3))
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
print("c5", round(d["c5_job"]["value"]), d["c5_job"]["hypotheses_checksum"], "cpu", round(d["cpu_baseline"]["value"], 3))
P
