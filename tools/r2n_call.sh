#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python tools/r2n_probe.py 2>&1 | tee gpurun_out/r2n_probe.log
