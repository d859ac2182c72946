#!/bin/bash
set -u
mkdir -p gpurun_out
for h in tcgen05 cublas tcgen05; do
CTCPS_HEAD=$h timeout 300 python bench.py --no-cpu-baseline --extra-configs "" --c5-utterances 0 --no-drop-in --pre-beam 0 > gpurun_out/r2o_bench_$h.json 2> gpurun_out/r2o_bench_$h.err; echo "bench $h rc=$?"
python - <<P
import json
d = json.loads(open("gpurun_out/r2o_bench_$h.json").read().strip().splitlines()[-1])
print("$h", {k: (round(v["value"]), round(v["ctc_head_ms"], 3), v["ctc_head_implementation"]) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict)})
P
done
