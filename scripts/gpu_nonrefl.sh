# non-reflecting BCs on the GPU: achieved errors for convectingVortex (printed), then the GPU suite
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/w.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, aither_b200, goldencheck as gc
mk = lambda prob: aither_b200.GridLevel(prob)
loose = dict(ghosts=1e-6, residual=1e-6, specRadius=1e-6, dt=1e-6, diag=1e-6, x0=1e-6, x=1e-6,
             matrixResid=1e-3, state=1e-6, l2=1e-6, turb=1e-6)
d = gc.load("convectingVortex")
try:
    print("convectingVortex phases", {k: float("%.2e" % v) for k, v in gc.check_phases(mk, d, 0, loose).items()}, flush=True)
    print("convectingVortex history", gc.check_history(mk, d, 40, 1e-6), flush=True)
except Exception as e:
    print("FAILED", repr(e)[:800], flush=True)
PY
timeout 300 python /tmp/w.py > gpurun_out/r01o_nonreflecting_errors.txt 2>&1; cat gpurun_out/r01o_nonreflecting_errors.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r01o_pytest_gpu.txt
