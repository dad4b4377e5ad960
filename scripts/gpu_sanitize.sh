cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/t.py <<'PY'
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, aither_b200
from aither_b200 import synthetic
prob = synthetic.box_problem(40, 20, 6, sweeps=2)
lvl = aither_b200.GridLevel(prob)
lvl.store_old_solution(0)
print(lvl.iterate(30.0))
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/t.py 2>&1 | head -60
