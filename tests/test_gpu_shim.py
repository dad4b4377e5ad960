"""The drop-in seam, compiled and run: oracle/_ref/aither_gpu_main is the reference's own main
program (its unmodified objects: input parser, Plot3D reader, decomposition, multigrid coarsening,
k-d tree wall distance, residual log, output writers) with the iteration body replaced by the
shim shim/gpuPath.hpp -> libaither_b200.so. It runs the reference's SHIPPED input files
(tests/cases/<name>/: .inp + grid as shipped in testCases/) exactly as
testCases/regressionTests.py runs the reference -- iterations edited to the suite's count, one
rank -- and the last line of the .resid file it writes must meet the reference's regression
goldens (1 %, the suite's own tolerance and ignore indices)."""
import os
import shutil
import subprocess
import tempfile

import pytest

import refcase

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "oracle", "_ref", "aither_gpu_main")

# name, iterations, regression golden (None = index the suite ignores), source line
CASES = [
    ("subsonicCylinder", 100, [1.8751e-01, 2.6727e-01, 3.1217e-01, None, 1.8639e-01]),    # :241
    ("multiblockCylinder", 100, [2.0529e-01, 3.4540e-01, 5.0153e-01, None, 1.9997e-01]),  # :260
    ("transonicBump", 100, [2.6152e-02, 1.5984e-02, 9.6803e-03, None, 1.9215e-02]),       # :333, 3-level W cycle
    ("viscousFlatPlate", 100, [7.4673e-02, 2.4711e-01, 3.8960e-02, None, 7.7683e-02]),    # :356
    ("turbFlatPlate", 20, [2.2309e-01, 2.9862e-01, None, 3.2376e-01, 2.1910e-01, 2.5208e-07,
                           3.3009e-06]),                                                  # :379
]


@pytest.mark.parametrize("name,iters,gold", [c for c in CASES if c[0] in ("viscousFlatPlate", "turbFlatPlate")],
                         ids=["viscousFlatPlate", "turbFlatPlate"])
def test_shim_with_the_wall_distance_from_the_device(name, iters, gold):
    """AITHER_GPU_WALLDIST=1: the shim skips the reference's k-d tree search and calls
    aither_gpu_compute_wall_distance; the regression goldens are met all the same"""
    test_shipped_case_through_the_shim(name, iters, gold, env_extra={"AITHER_GPU_WALLDIST": "1"},
                                       expect="wall distance from")


@pytest.mark.parametrize("name,iters,gold", CASES, ids=[c[0] for c in CASES])
def test_shipped_case_through_the_shim(name, iters, gold, env_extra=None, expect=None):
    if not os.path.exists(BINARY):
        pytest.skip("oracle/_ref/aither_gpu_main has not been built (make -C oracle shim)")
    with tempfile.TemporaryDirectory() as tmp:
        inp = refcase.stage_case(os.path.join(ROOT, "tests", "cases", name), tmp,
                                 edits={"outputFrequency": str(iters)}, iterations=iters)
        with open(os.path.join(tmp, "air.dat"), "w") as f:
            f.write(refcase.AIR_DAT)
        env = dict(os.environ, AITHER_INSTALL_DIRECTORY=tmp, **(env_extra or {}))
        res = subprocess.run([BINARY, inp], cwd=tmp, env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-3000:]
        assert "gpuPath: aither_b200" in res.stdout and "Program Complete" in res.stdout
        assert expect is None or expect in res.stdout, res.stdout[-3000:]
        last = open(os.path.join(tmp, name + ".resid")).readlines()[-1].split()
        # same columns as regressionTests.py GetTestCaseResiduals
        mine = [float(v) for v in last[3:3 + len(gold)]]
        for e, gv in enumerate(gold):
            if gv is not None:
                assert abs(mine[e] - gv) <= 0.01 * gv, (name, e, mine, gold)
        # the function file of the last iteration was written from the downloaded state
        assert any(f.endswith(".fun") for f in os.listdir(tmp)), os.listdir(tmp)
