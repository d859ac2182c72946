#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): hardware gate of the time-parallel select kernel, then A/B, then profiles.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Everything lands in gpurun_out/r2a_*; nothing here is a bench value except the two bench.py lines at the end.
set -u
mkdir -p gpurun_out
# 1. gate: sequential == time-parallel (parity criterion), golden replays, pre-beam, native-loop 1-best
CTCPS_TEST_PSCAN=1 timeout 300 python -m pytest tests/test_gpu_select_pscan.py -q -m gpu > gpurun_out/r2a_pscan_tests.log 2>&1
echo "pscan tests rc=$?"; tail -3 gpurun_out/r2a_pscan_tests.log
# 2. A/B of the select call (stage + scan), graph-replayed, on the four shapes
for c in C1 C3 C4 C2; do for m in 0 1; do
  echo "== $c select_pscan=$m"; timeout 120 python tools/kernel_bench.py --config $c --only select --select-pscan $m 2>&1 | grep -i "select lazy"
done; done > gpurun_out/r2a_select_ab.log 2>&1
paste - - < gpurun_out/r2a_select_ab.log
# 3. whole decodes with the time-parallel selection (compare with profiles/r1x_bench.json, r1y_bench_C1.json)
CTCPS_SELECT_PSCAN=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_pscan.json 2> gpurun_out/r2a_bench_pscan.err; echo "bench C2 rc=$?"
CTCPS_SELECT_PSCAN=1 timeout 120 python bench.py --config C1 --no-cpu-baseline > gpurun_out/r2a_bench_pscan_C1.json 2> gpurun_out/r2a_bench_pscan_C1.err; echo "bench C1 rc=$?"
python - <<'P'
import json
for f in ("r2a_bench_pscan", "r2a_bench_pscan_C1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["pre_beam"]["value"]), d["roofline"]["avg_launch_ms"])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
P
