#!/bin/bash
# r3g: is the C5 job really slower with the frame window on?  alternate the two settings on one box; then the new bench line
set -u
mkdir -p gpurun_out
for rep in 1 2; do for on in 0 1; do
CTCPS_FRAME_WINDOW=$on timeout 300 python bench.py --no-cpu-baseline --no-drop-in --pre-beam 0 --hidden-dim 0 --extra-configs "" --steps 3 --warmup 3 > gpurun_out/r3g_c5_win${on}_$rep.json 2> gpurun_out/r3g_c5_win${on}_$rep.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3g_c5_win${on}_$rep.json").read().strip().splitlines()[-1])
print("rep $rep frame_window=$on:", "C2", round(d["value"]), "c5", round(d["c5_job"]["value"]), round(d["c5_job"]["ms"], 1), d["c5_job"]["hypotheses_checksum"])
P
done; done
( time timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err ) 2>&1 | grep real
python - <<P
import json
d = json.loads(open("gpurun_out/r3g_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline achieved", round(r["achieved"]), "frac", round(r["frac"], 3), "avg ms", round(r["avg_launch_ms"], 4), "streamed frac", round(r.get("streamed_fraction_of_all_frames", 0), 3), "launches", d["gpu_launches"])
e = d["every_row_every_frame"]
print("every row every frame:", round(e["value"]), "frac", round(e["roofline"]["frac"], 3), "ms", round(e["roofline"]["avg_launch_ms"], 4))
print("configs", {k: (round(v["value"]), round(v["roofline"]["frac"], 3), round(v["roofline"].get("streamed_fraction_of_all_frames", 0), 2)) for k, v in d["configs"].items()})
print("hidden", {k: round(v["value"]) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict) and "value" in v})
for k, v in d.get("drop_in", {}).items():
    print("drop_in", k, v.get("unavailable") or (round(v["value"]), round(v["e2e"]["value"]), v["transcripts_recovered"]))
print("c5", round(d["c5_job"]["value"]), d["c5_job"]["hypotheses_checksum"])
P
