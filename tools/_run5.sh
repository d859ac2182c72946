timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/r1x_tests_full.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r1x_tests_full.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1x_smoke.log
timeout 400 python bench.py > gpurun_out/r1x_bench.json 2> gpurun_out/r1x_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r1x_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'])
P
