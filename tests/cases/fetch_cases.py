"""Copy the INPUT DATA (no source code) of the reference's shipped test cases that
tests/test_gpu_shim.py runs -- the .inp file and the Plot3D grid, read-only data of
mnucci32/aither v0.10.0 testCases/ -- into tests/cases/<name>/, so that the tests can run on the
GPU box where /root/reference does not exist. Run here: python tests/cases/fetch_cases.py"""
import os
import shutil

REF = "/root/reference/testCases"
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["subsonicCylinder", "multiblockCylinder", "turbFlatPlate", "transonicBump",
         "viscousFlatPlate"]

for name in CASES:
    dst = os.path.join(HERE, name)
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(os.path.join(REF, name)):
        if f.endswith((".inp", ".xyz")):
            shutil.copy(os.path.join(REF, name, f), os.path.join(dst, f))
            os.chmod(os.path.join(dst, f), 0o644)
    print(name, sorted(os.listdir(dst)))
