# round 2, step z: LU-SGS pencil kernel with the fence-free mailbox hand-over: parity, time lines, timing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_phases.py tests/test_gpu_multiblock.py tests/test_gpu_rans.py tests/test_gpu_viscous.py -m gpu -q -x -k "lusgs or subsonicCylinder or viscousFlatPlate or turbFlatPlate or box_kw or uniformFlow or inlet_outlet or periodic or multiblock or shockTube or LUSGS or Lusgs" > gpurun_out/r02z_pytest_lusgs.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02z_pytest_lusgs.txt
cat > /tmp/tl.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import aither_b200
from aither_b200 import synthetic
ni, nj, nk = [int(v) for v in sys.argv[1:4]]
prob = synthetic.box_problem(ni, nj, nk, solver="lusgs", sweeps=2)
gl = aither_b200.GridLevel(prob, device=0)
for it in range(4):
    gl.store_old_solution(it); gl.iterate(50.0)
gl.close()
PY
tl() { AITHER_B200_LUSGS_DBG=gpurun_out/r02z_timeline_$1.txt timeout 120 python /tmp/tl.py $2 $3 $4; }
tl 128x8x64 128 8 64
tl 128x128x8 128 128 8
tl 128x64x64 128 64 64
tl 192 192 192 192
timeout 300 python scripts/probe_lusgs.py > gpurun_out/r02z_lusgs_probe.json 2> gpurun_out/r02z_probe.err; echo "probe rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --n 192 --solver lusgs > gpurun_out/r02z_lusgs192.json 2> gpurun_out/r02z_lusgs192.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02z_lusgs_probe.json'):
    r=json.loads(l); print(r['shape'], 'half sweep %.3f ms'%r['ms_per_half_sweep'], 'us/plane %.3f'%r['us_per_plane'])
d=json.loads([l for l in open('gpurun_out/r02z_lusgs192.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['kernel_ms_per_step'])
PY
