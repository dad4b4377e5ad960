# round 2, step t: state update fused into the matrix-residual pass (alternate state buffer), coalesced BC kernel; headline bench A/B; whole suite
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02t_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError|passed|failed" gpurun_out/r02t_pytest_gpu.txt | head -30
run() { name=$1; shift; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu "$@" > gpurun_out/r02t_$name.json 2> gpurun_out/r02t_$name.err || tail -3 gpurun_out/r02t_$name.err; }
run base
AITHER_B200_FUSE_UPDATE=0 run unfused_update
run lusgs192 --n 192 --solver lusgs
run sst_blusgs96 --n 96 --turb sst2003 --solver blusgs
run sst_lusgs128 --n 128 --turb sst2003 --solver lusgs
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02t_')[1][:-5], 'ms/step %.3f' % d['ms_per_step'], 'iter roofline %.3f'%d['roofline_iteration']['frac'], d['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(f, 'failed', e)
PY
