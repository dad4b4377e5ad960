# round 2, step aa: whole GPU suite with the mailbox LU-SGS kernel; headline + LU-SGS / RANS variants
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02aa_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError|passed|failed" gpurun_out/r02aa_pytest_gpu.txt | head -30
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/r02aa_$name.json 2> gpurun_out/r02aa_$name.err || tail -3 gpurun_out/r02aa_$name.err; }
run lusgs192 --n 192 --solver lusgs
run sst_lusgs128 --n 128 --turb sst2003 --solver lusgs
run sst_blusgs96 --n 96 --turb sst2003 --solver blusgs
run visc_lusgs128 --n 128 --viscous --solver lusgs
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aa_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02aa_')[1][:-5], 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
