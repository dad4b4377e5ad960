# round 2, step v: wall-data / output-staging tests; LU-SGS pencil cross-section A/B (192^3)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "wall_data or output or shim or golden" > gpurun_out/r02v_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError|passed|failed" gpurun_out/r02v_pytest_gpu.txt | head -20
run() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --n 192 --solver lusgs > gpurun_out/r02v_$name.json 2> gpurun_out/r02v_$name.err || tail -3 gpurun_out/r02v_$name.err; }
run p8x7x1
for v in p4x8x2 p8x4x2 p8x8x1; do AITHER_B200_LIB=$PWD/aither_b200/lib/variants/lib_$v.so run $v; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02v_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02v_')[1][:-5], 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
