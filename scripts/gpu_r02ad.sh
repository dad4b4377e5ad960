# round 2, step ad: cross-section of the one-thread-per-cell pencil kernel (192^3 LU-SGS x4; SST 128^3)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { name=$1; shift; timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/r02ad_$name.json 2> gpurun_out/r02ad_$name.err || tail -3 gpurun_out/r02ad_$name.err; }
run p16x8 --n 192 --solver lusgs
for v in p12x8 p16x6 p8x8 p8x12; do AITHER_B200_LIB=$PWD/aither_b200/lib/variants/lib_$v.so run $v --n 192 --solver lusgs; done
run sst_p16x8 --n 128 --turb sst2003 --solver lusgs
for v in p12x8 p8x8; do AITHER_B200_LIB=$PWD/aither_b200/lib/variants/lib_$v.so run sst_$v --n 128 --turb sst2003 --solver lusgs; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ad_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02ad_')[1][:-5], 'ms/step %.3f' % d['ms_per_step'], 'wave %.3f'%d['kernel_ms_per_step']['lusgs_plane'])
    except Exception as e:
        print(f, 'failed', e)
PY
