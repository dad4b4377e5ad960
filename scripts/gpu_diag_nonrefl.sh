# diagnostic: Euler run with non-reflecting BCs, GPU against the oracle phase by phase over two iterations
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/w.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, aither_b200, goldencheck as gc, oracle, refcase
from aither_b200 import ctypes_abi as abi
d = gc.load("box_nonrefl_euler")
prob = refcase.problem_from_dump(d, state_key="state0")
gpu, orc = aither_b200.GridLevel(prob), oracle.OracleLevel(prob)
def cmp(tag, fld):
    a, b = gpu.field(0, fld), orc.field(0, fld)
    err = np.abs(a - b)
    sc = np.abs(b).max()
    w = np.unravel_index(err.argmax(), err.shape)
    print("%-28s max|b| %.3e  rel err %.3e at %s gpu %.6e orc %.6e" % (tag, sc, err.max() / max(sc, 1e-300), w, a[w], b[w]), flush=True)
for it in range(2):
    cfl = float(d["hist/cfl"][it])
    for l in (gpu, orc):
        l.store_old_solution(it)
        l.get_boundary_conditions()
    cmp("it%d state after BC" % it, abi.FIELD_STATE)
    for l in (gpu, orc):
        l.calc_residual()
    cmp("it%d residual" % it, abi.FIELD_RESIDUAL)
    cmp("it%d velGrad" % it, abi.FIELD_VELOCITY_GRAD)
    cmp("it%d pressGrad" % it, abi.FIELD_PRESSURE_GRAD)
    for l in (gpu, orc):
        l.calc_time_step(cfl); l.invert_diagonal(); l.initialize_matrix_update()
    cmp("it%d dt" % it, abi.FIELD_DT)
    for l in (gpu, orc):
        l.relax(); l.update_blocks()
    cmp("it%d state after update" % it, abi.FIELD_STATE)
    cmp("it%d consN" % it, abi.FIELD_CONS_N)
PY
timeout 300 python /tmp/w.py 2>&1 | tail -30 | tee gpurun_out/diag_nonrefl.txt
