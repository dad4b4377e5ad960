# round 2, step m: hoisted classification, butterfly reduction; ahead kernel timed on its own
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python scripts/probe_lusgs.py gpurun_out/r02m_probe.json 2>&1 | tail -12
AITHER_B200_LUSGS_DBG=1 timeout 300 python - 512 8 7 2> gpurun_out/r02m_stamps_lone.txt <<'PY'
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
head -8 gpurun_out/r02m_stamps_lone.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or phases or rans or multiblock" > gpurun_out/r02m_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02m_pytest_gpu.txt
