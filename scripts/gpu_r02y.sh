# round 2, step y: time line of the pencils of one forward half sweep; experiments with parts of the hand-over switched off
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_golden.py tests/test_gpu_multiblock.py -m gpu -q -x -k "subsonicCylinder or multiblock or box_lusgs" 2>&1 | tail -2
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
tl() { AITHER_B200_LUSGS_DBG=gpurun_out/r02y_timeline_$1.txt timeout 120 python /tmp/tl.py $2 $3 $4; }
tl 128x64x64 128 64 64
tl 192 192 192 192
AITHER_B200_LUSGS_DBGFLAGS=4 tl 128x64x64_nowait 128 64 64
