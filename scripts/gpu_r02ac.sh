# round 2, step ac: interior-only upload (e2e), whole GPU suite of the final build, default bench
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ac_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError|passed|failed" gpurun_out/r02ac_pytest_gpu.txt | head -20
timeout 600 python bench.py > gpurun_out/r02ac_bench.json 2> gpurun_out/r02ac_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ac_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], {k:d['e2e'][k] for k in ('value','h2d_bytes_per_step','ms_per_step','mode','value_overlapped','value_synchronous')}, d['roofline']['frac'], d['roofline']['traffic'], d['kernel_ms_per_step'])
PY
