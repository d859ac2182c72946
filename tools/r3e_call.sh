#!/bin/bash
# r3e: warp-per-survivor ranking in k_beam_merge: beam-step tests, the whole GPU suite, timelines of C4 / C2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r3e_tests.log; cat gpurun_out/r3e_tests.log
for c in C4 C2; do
timeout 300 python tools/timeline.py --config $c --tag r3e_$c > gpurun_out/r3e_tl_$c.log 2>&1; echo "timeline $c rc=$?"; sed -n 2,12p gpurun_out/r3e_${c}_timeline.txt
done
