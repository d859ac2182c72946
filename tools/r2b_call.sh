#!/bin/bash
# r2b: new fused top-k path, whole GPU suite, bench A/B (fused vs unfused step)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_topk.py -q -m gpu -x > gpurun_out/r2b_fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -5 gpurun_out/r2b_fused_tests.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2b_tests_full.log 2>&1
echo "full tests rc=$?"; tail -8 gpurun_out/r2b_tests_full.log
timeout 300 python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 > gpurun_out/r2b_bench_fused.json 2> gpurun_out/r2b_bench_fused.err; echo "bench fused rc=$?"
timeout 300 python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 --no-fuse-topk > gpurun_out/r2b_bench_unfused.json 2> gpurun_out/r2b_bench_unfused.err; echo "bench unfused rc=$?"
for c in C1 C3 C4; do
timeout 300 python bench.py --config $c --no-cpu-baseline --single-mode --hidden-dim 0 > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err; echo "bench $c rc=$?"
done
python - <<'P'
import json
for f in ("r2b_bench_fused", "r2b_bench_unfused", "r2b_bench_C1", "r2b_bench_C3", "r2b_bench_C4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["avg_launch_ms"], round(d["roofline"]["frac"], 3), d["config"]["decode_steps_per_utterance_batch"])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
P
