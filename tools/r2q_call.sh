#!/bin/bash
# r2q: GPU tests of the tree with the window parameters threaded through the kernels, then device timelines (CUPTI) of the native loop
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2q_tests.log; cat gpurun_out/r2q_tests.log
for c in C2 C4 C1 C3; do
timeout 300 python tools/timeline.py --config $c --tag r2q_$c > gpurun_out/r2q_tl_$c.log 2>&1; echo "timeline $c rc=$?"
done
timeout 300 python tools/timeline.py --config C2 --state pre_beam --tag r2q_C2pre > gpurun_out/r2q_tl_C2pre.log 2>&1; echo "timeline C2 pre rc=$?"
cat gpurun_out/r2q_C2_timeline.txt
