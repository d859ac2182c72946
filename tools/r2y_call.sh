#!/bin/bash
# r2y: GPU tests (new: head scale test), smoke(), short bench with the head roofline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2y_tests.log; cat gpurun_out/r2y_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2y_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "score_ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
h = d["e2e_from_hidden"]
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3)) for k, v in h.items() if isinstance(v, dict) and "value" in v})
print("head_roofline", {k: h["head_roofline"][k] for k in ("achieved", "peak", "frac", "avg_launch_ms")})
P
