# compute-sanitizer racecheck (shared-memory hazards) over small LU-SGS runs: the pencil kernel hands
# ingredients between threads through shared memory with one barrier per plane
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/r.py <<'PY'
import sys
sys.path.insert(0,'.')
import aither_b200
from aither_b200 import synthetic
def run(prob, cfl=20.0, n=2):
    lvl = aither_b200.GridLevel(prob)
    for it in range(n):
        lvl.store_old_solution(it)
        out = lvl.iterate(cfl)
    lvl.close()
    return out[0]
print("euler lusgs, 3 x 3 pencils", run(synthetic.box_problem(24, 30, 20, solver="lusgs", sweeps=2)))
print("laminar lusgs, 2 x 2 pencils", run(synthetic.box_problem(16, 20, 12, solver="lusgs", viscous=True, size=2e-5, sweeps=2)))
print("euler dplur (TMA sweep)", run(synthetic.box_problem(40, 20, 18, sweeps=2)))
PY
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python /tmp/r.py > gpurun_out/r02_racecheck.log 2>&1
grep -c "hazard" gpurun_out/r02_racecheck.log; tail -12 gpurun_out/r02_racecheck.log
