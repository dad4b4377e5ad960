# secondary workloads (not the headline): kernel times of the WENO / viscous / RANS / block-matrix paths
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || tail -3 gpurun_out/var_$name.err; }
run weno192 --n 192 --recon weno
run visc192 --n 192 --recon weno --viscous
run sst128 --n 128 --turb sst2003
run kw128 --n 128 --turb kOmegaWilcox2006
run bdplur128 --n 128 --solver bdplur
run sst_blusgs96 --n 96 --turb sst2003 --solver blusgs
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/var_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('var_')[1][:-5], 'ms/step %.2f' % d['ms_per_step'], 'Mcell-iter/s %.0f' % d['value'], d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
