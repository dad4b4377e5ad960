"""Generate the wall-distance fixtures from the UNMODIFIED reference (oracle/_ref/aither_dump).

TEST INFRASTRUCTURE ONLY. Run in the build container, where /root/reference exists:

    python tests/golden/make_walldist.py

One compressed .npz per case with what the reference's set-up knows about the wall distance:
block dimensions, boundary surfaces, cell centres, the face centres of the three directions
(from which the tests collect the viscous-wall face centres as GetViscousFaceCenters does,
src/utility.cpp:310-368) and the reference's wallDist_ array, ghost cells included
(procBlock::CalcWallDistance, src/procBlock.cpp:6030-6107, k-d tree src/kdtree.cpp).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refcase  # noqa: E402

REF_CASES = "/root/reference/testCases"
# laminar plate (one block, the wall covers part of a j-face), the wall-law case (two blocks,
# SST), couette (moving wall + periodic pair)
CASES = ("viscousFlatPlate", "wallLaw", "couette")
KEEP = ("dims", "surfaces", "center", "fCenterI", "fCenterJ", "fCenterK", "wallDist")


def generate(name):
    with tempfile.TemporaryDirectory() as tmp:
        inp = refcase.stage_case(os.path.join(REF_CASES, name), tmp, {}, iterations=1)
        d = refcase.run_harness(tmp, inp, 1, full=(), geom=True)
    out = {k: np.asarray(v) for k, v in d.items()
           if k == "numBlocks" or k == "cfg/numGhosts" or k.split("/")[-1] in KEEP}
    path = os.path.join(HERE, "walldist_" + name + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %d records, %.1f kB" % (name, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    if not refcase.have_harness():
        sys.exit("oracle/_ref/aither_dump is not built (make -C oracle ref)")
    for nm in (sys.argv[1:] or CASES):
        generate(nm)
