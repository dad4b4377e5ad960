# round 2, step y: time line of the pencils of one forward half sweep
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
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
tl 128x8x64 128 8 64
tl 128x128x8 128 128 8
tl 128x64x64 128 64 64
tl 192 192 192 192
