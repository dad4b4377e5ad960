# round 2, step g: ncu of a LONE pencil (512 x 8 x 8 block: one thread block, no waiting on neighbours)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/lone.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import aither_b200
from aither_b200 import synthetic
prob = synthetic.box_problem(512, 8, 8, solver="lusgs", sweeps=2)
gl = aither_b200.GridLevel(prob, device=0)
for it in range(2):
    gl.store_old_solution(it); gl.iterate(50.0)
gl.close()
PY
timeout 600 ncu --set full --warp-sampling-interval 0 --warp-sampling-buffer-size 536870912 --clock-control none --import-source on -k regex:LusgsPencil -s 6 -c 2 -o gpurun_out/r02g_lone -f python /tmp/lone.py > gpurun_out/r02g_ncu.log 2>&1; tail -2 gpurun_out/r02g_ncu.log
