#!/bin/bash
# r3c: eight ranks of one box: the full bench line under torchrun (weak scaling of the step + the C5 job sharded over 8 GPUs)
set -u
mkdir -p gpurun_out
N=${1:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r3c_bench_n$N.json 2> gpurun_out/r3c_bench_n$N.err ) 2>&1 | grep real; echo "bench n$N rc=$?"
tail -2 gpurun_out/r3c_bench_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/r3c_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N C2", round(d["value"]), "e2e", round(d["e2e"]["value"]), "n_gpus", d["n_gpus"], "scaling", d["scaling"])
print("hidden", {k: (round(v["value"]), round(v["ctc_head_ms"], 3)) for k, v in d["e2e_from_hidden"].items() if isinstance(v, dict) and "value" in v})
print("c5", {k: d["c5_job"].get(k) for k in ("value", "ms", "n_gpus", "utterances_differing_from_aligned_transcript", "copies_of_an_utterance_agree", "hypotheses_checksum")})
P
