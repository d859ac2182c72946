#!/bin/bash
# r2d: correctness of the reworked pipeline, then A/B sweeps: L2 look-ahead depth, hypothesis-group width
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2d_tests_full.log 2>&1
echo "full tests rc=$?"; tail -4 gpurun_out/r2d_tests_full.log
B="python bench.py --no-cpu-baseline --single-mode --hidden-dim 0 --steps 4"
for pf in 0 2 4 8; do
  CTCPS_PSI_PREFETCH=$pf timeout 200 $B > gpurun_out/r2d_C2_pf$pf.json 2> gpurun_out/r2d_C2_pf$pf.err; echo "C2 pf$pf rc=$?"
done
CTCPS_PSI_PREFETCH=2 timeout 200 $B --no-fuse-topk > gpurun_out/r2d_C2_unfused_pf2.json 2> gpurun_out/r2d_C2_unfused.err; echo "C2 unfused rc=$?"
for g in 10 20; do for pf in 0 4; do
  CTCPS_PSI_MAX_GROUP=$g CTCPS_PSI_PREFETCH=$pf timeout 200 $B --config C4 > gpurun_out/r2d_C4_g${g}_pf$pf.json 2> gpurun_out/r2d_C4_g${g}_pf$pf.err; echo "C4 g$g pf$pf rc=$?"
done; done
for c in C1 C3; do for pf in 0 4; do
  CTCPS_PSI_PREFETCH=$pf timeout 200 $B --config $c > gpurun_out/r2d_${c}_pf$pf.json 2> gpurun_out/r2d_${c}_pf$pf.err; echo "$c pf$pf rc=$?"
done; done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{f[11:-5]:24s} utt/s {d['value']:9.0f}  ms/step {d['ms_per_step']:8.3f}  score_ms {d['roofline']['avg_launch_ms']:.4f}  frac {d['roofline']['frac']:.3f}  launches {d['gpu_launches']}")
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
P
