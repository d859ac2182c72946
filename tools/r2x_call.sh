#!/bin/bash
# r2x: two ranks of one box: the full bench line under torchrun (weak scaling of the step + the C5 job sharded over 2 GPUs)
set -u
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err ) 2>&1 | grep real; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2x_bench_n2.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2x_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], "scaling", d["scaling"])
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3)) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict) and "value" in v})
print("c5", {k: d["c5_job"].get(k) for k in ("value", "ms", "ranks", "utterances_differing_from_aligned_transcript", "copies_of_an_utterance_agree", "hypotheses_checksum")})
P
