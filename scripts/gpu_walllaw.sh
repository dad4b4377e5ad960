# wall-law parity on the GPU: achieved per-phase errors of the three wall-law fixtures (printed, so
# that the bars in the tests can be set from measurements), then the whole GPU suite, the default
# bench line and the SST variant (the RANS cell kernel gained the wall-law branch)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
cat > /tmp/w.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, aither_b200, goldencheck as gc
mk = lambda prob: aither_b200.GridLevel(prob)
loose = dict(ghosts=1e-6, residual=1e-6, specRadius=1e-6, dt=1e-6, diag=1e-6, x0=1e-6, x=1e-6,
             matrixResid=1e-3, state=1e-6, l2=1e-6, turb=1e-6)
for nm, n in (("wallLaw_cloud", 6), ("box_walllaw_isothermal", 10), ("box_walllaw_heatflux", 10), ("wallLaw", 20)):
    d = gc.load(nm)
    try:
        if nm != "wallLaw":
            print(nm, "phases", {k: float("%.2e" % v) for k, v in gc.check_phases(mk, d, 0, loose).items()}, flush=True)
        print(nm, "history", gc.check_history(mk, d, n, 1e-6), flush=True)
    except Exception as e:
        print(nm, "FAILED", repr(e)[:600], flush=True)
PY
timeout 600 python /tmp/w.py > gpurun_out/r01n_walllaw_errors.txt 2>&1; cat gpurun_out/r01n_walllaw_errors.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r01n_pytest_gpu.txt
timeout 300 python bench.py > gpurun_out/r01n_bench.json 2> gpurun_out/r01n_bench.err; tail -c 1500 gpurun_out/r01n_bench.json
timeout 300 python bench.py --n 128 --turb sst2003 > gpurun_out/r01n_variant_sst.json 2>/dev/null; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r01n_variant_sst.json").read().strip().splitlines()[-1])
print("sst 128^3:", d["ms_per_step"], d["kernel_ms_per_step"])
PY
