# round 2, step l: clock64 stamps inside one pencil (where does a plane's time go?)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/lone.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import aither_b200
from aither_b200 import synthetic
shape = tuple(int(v) for v in sys.argv[1:4])
prob = synthetic.box_problem(*shape, solver="lusgs", sweeps=2)
gl = aither_b200.GridLevel(prob, device=0)
for it in range(4):
    gl.store_old_solution(it); gl.iterate(50.0)
gl.close()
PY
AITHER_B200_LUSGS_DBG=1 timeout 300 python /tmp/lone.py 512 8 7 2> gpurun_out/r02l_stamps_lone.txt; head -40 gpurun_out/r02l_stamps_lone.txt
AITHER_B200_LUSGS_DBG=1 timeout 300 python /tmp/lone.py 192 192 192 2> gpurun_out/r02l_stamps_192.txt; head -40 gpurun_out/r02l_stamps_192.txt
