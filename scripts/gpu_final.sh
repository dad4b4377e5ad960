# final check of the round: smoke() and the default bench line without the CPU sample
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 200 python bench.py --no-cpu > gpurun_out/r01q_bench_nocpu.json 2>gpurun_out/r01q_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r01q_bench_nocpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline_iteration"]["frac"], d["e2e"]["value"], d["kernel_ms_per_step"])
PY
