"""Which restructured operation costs the residual its last digit? Runs goldencheck.check_phases
(reference dumps, every component against its own scale) for the given fixtures / iterations with
the shipped library and with each bisect build (scripts/build_bisect.sh: one operation switched back
to the reference's form), one subprocess per library, and prints the achieved per-phase errors.
Test infrastructure; run on the GPU box: python scripts/diag_bisect.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("subsonicCylinder", None), ("multiblockCylinder", None), ("turbFlatPlate", None)]

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import goldencheck as gc
    import aither_b200
    big = dict(ghosts=1, residual=1, specRadius=1, dt=1, diag=1, x0=1, x=1, matrixResid=1e9, state=1,
               l2=1, turb=1)
    res = {}
    for name, _ in CASES:
        d = gc.load(name)
        for it in gc.full_iterations(d):
            out = gc.check_phases(aither_b200.GridLevel, d, it, big)
            res["%s it%d" % (name, it)] = {k: out[k] for k in ("ghosts", "residual", "x", "state", "l2")}
    print("RESULT " + json.dumps(res))
    sys.exit(0)

libs = [("shipped", None)] + [(v, os.path.join(ROOT, "aither_b200", "lib", "bisect", "lib_%s.so" % v))
                              for v in ("EXACT_RCP", "REF_ROE", "MUSCL_DIV")]
table = {}
for label, path in libs:
    if path is not None and not os.path.exists(path):
        continue
    env = dict(os.environ)
    if path:
        env["AITHER_B200_LIB"] = path
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    if not line:
        print(label, "FAILED\n", r.stdout[-2000:])
        continue
    table[label] = json.loads(line[0][7:])
keys = sorted(next(iter(table.values())).keys())
print("%-28s" % "per-cell residual rel. err" + "".join("%12s" % k for k in table))
for k in keys:
    print("%-28s" % k + "".join("%12.2e" % table[lab][k]["residual"] for lab in table))
print("%-28s" % "ghost cells after BCs")
for k in keys:
    print("%-28s" % k + "".join("%12.2e" % table[lab][k]["ghosts"] for lab in table))
