# round 2, step b: pencil wavefront LU-SGS with shared ingredients -- parity (whole -m gpu suite) and timing vs the per-plane launches
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r02b_pytest_gpu.txt
run() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err || tail -3 gpurun_out/r02b_$name.err; }
run lusgs192 --n 192 --solver lusgs
AITHER_B200_LUSGS=planes run lusgs192_planes --n 192 --solver lusgs
run lusgs256 --n 256 --solver lusgs
run sst_blusgs96 --n 96 --turb sst2003 --solver blusgs
run sst_lusgs128 --n 128 --turb sst2003 --solver lusgs
run visc_lusgs128 --n 128 --viscous --solver lusgs
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('r02b_')[1][:-5], 'ms/step %.2f' % d['ms_per_step'], 'Mcell-iter/s %.0f' % d['value'], d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
