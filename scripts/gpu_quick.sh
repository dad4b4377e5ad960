# quick GPU check: all -m gpu tests, then the default bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -3 gpurun_out/quick_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['mode'], d['e2e']['value_overlapped'], d['e2e']['value_synchronous'])
print('roofline', d['roofline']['kernel'], d['roofline']['frac'], 'iter frac', d['roofline_iteration']['frac'])
print(d['kernel_ms_per_step'])
PY
