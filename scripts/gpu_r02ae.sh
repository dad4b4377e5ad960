# round 2, step ae: LU-SGS records without the viscous slots for Euler runs, 12 x 8 pencils: parity + timing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_phases.py tests/test_gpu_multiblock.py tests/test_gpu_rans.py tests/test_gpu_viscous.py -m gpu -q -x -k "lusgs or subsonicCylinder or viscousFlatPlate or turbFlatPlate or box_kw or uniformFlow or inlet_outlet or periodic or multiblock or shockTube or LUSGS or Lusgs" > gpurun_out/r02ae_pytest_lusgs.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02ae_pytest_lusgs.txt
run() { name=$1; shift; timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/r02ae_$name.json 2> gpurun_out/r02ae_$name.err || tail -3 gpurun_out/r02ae_$name.err; }
run lusgs192 --n 192 --solver lusgs
run sst_lusgs128 --n 128 --turb sst2003 --solver lusgs
run visc_lusgs128 --n 128 --viscous --solver lusgs
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ae_*.json')):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f.split('r02ae_')[1][:-5], 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms_per_step'])
PY
