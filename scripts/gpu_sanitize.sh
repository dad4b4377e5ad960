# compute-sanitizer memcheck over one small run of every kernel family: Euler (TMA sweep, fused
# prep, prefetch), LU-SGS (graph + split kernel), laminar, RANS, block matrices, three species,
# multi-block + periodic exchange
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/t.py <<'PY'
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, aither_b200
import goldencheck as gc, refcase
from aither_b200 import synthetic
def run(prob, cfl=20.0, n=2):
    lvl = aither_b200.GridLevel(prob)
    for it in range(n):
        lvl.store_old_solution(it)
        out = lvl.iterate(cfl)
    lvl.close()
    return out[0]
print("euler dplur", run(synthetic.box_problem(40, 20, 18, sweeps=2)))
print("euler weno lusgs", run(synthetic.box_problem(20, 12, 10, solver="lusgs", sweeps=2, recon="weno")))
print("euler lusgs, 3 x 3 pencils", run(synthetic.box_problem(24, 30, 20, solver="lusgs", sweeps=2)))
print("laminar lusgs, 2 x 2 pencils", run(synthetic.box_problem(16, 20, 12, solver="lusgs", viscous=True, size=2e-5, sweeps=2)))
print("laminar", run(synthetic.box_problem(20, 12, 10, viscous=True, size=2e-5, sweeps=2)))
print("sst", run(synthetic.box_problem(14, 10, 9, turb="sst2003", limiter="vanAlbada", size=1e-3, sweeps=2), 5.0))
print("sst blusgs", run(synthetic.box_problem(12, 10, 9, turb="sst2003", solver="blusgs", limiter="vanAlbada", size=1e-3, sweeps=2), 5.0))
print("bdplur", run(synthetic.box_problem(16, 10, 9, solver="bdplur", sweeps=2)))
for name in ("box_mix3_sst", "box_periodic", "uniformFlow_rans"):
    d = gc.load(name)
    print(name, run(refcase.problem_from_dump(d, "state0"), float(d["hist/cfl"][0])))
sp = synthetic.split_problem(synthetic.box_problem(16, 12, 10, sweeps=2), (2, 2, 1))
print("split", run(sp))
PY
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/t.py > gpurun_out/sanitize.log 2>&1
grep -c "Invalid\|Error" gpurun_out/sanitize.log; tail -25 gpurun_out/sanitize.log
