# round 2, step d: relaxed flag polling -- quick parity (lusgs-related tests) + timing + ncu
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or phases or rans or multiblock" > gpurun_out/r02d_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02d_pytest_gpu.txt
run() { name=$1; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err || tail -3 gpurun_out/r02d_$name.err; }
run lusgs192 --n 192 --solver lusgs
run lusgs128 --n 128 --solver lusgs
run sst_blusgs96 --n 96 --turb sst2003 --solver blusgs
run visc_lusgs128 --n 128 --viscous --solver lusgs
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('r02d_')[1][:-5], 'ms/step %.2f' % d['ms_per_step'], 'Mcell-iter/s %.0f' % d['value'], d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:LusgsPencil -s 3 -c 1 -o gpurun_out/r02d_pencil -f python bench.py --n 128 --solver lusgs --steps 1 --warmup 1 --no-cpu > gpurun_out/r02d_ncu.log 2>&1; tail -1 gpurun_out/r02d_ncu.log
