"""Wall distance on the device (-m gpu): aither_gpu_compute_wall_distance against the reference's
own wallDist_ arrays (tests/golden/walldist_*.npz: k-d tree search + ghost-cell rule of
procBlock::CalcWallDistance, src/procBlock.cpp:6030-6107), for the shipped laminar plate, the
two-block wall-law case and couette. The level is created with a wall distance of zero
everywhere, so what is compared was computed on the device."""
import numpy as np
import pytest

import goldencheck as gc
import refcase
import walldist_ref as wr
from aither_b200 import ctypes_abi as abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["viscousFlatPlate", "wallLaw", "couette"])
def test_device_wall_distance_matches_reference(name):
    import aither_b200
    geo = wr.load(name)
    prob = refcase.problem_from_dump(gc.load(name))
    for b in prob.blocks:
        b.arrays["wallDist"] = np.zeros_like(b.arrays["wallDist"])
    gpu = aither_b200.GridLevel(prob)
    pts = wr.wall_face_centers(geo, abi.BC_VISCOUS_WALL)
    gpu.compute_wall_distance(pts)
    g = prob.cfg.numGhosts
    for bb in range(len(prob.blocks)):
        ref = geo["b%d/wallDist" % bb][..., 0]
        mine = gpu.field(bb, abi.FIELD_WALL_DIST)[..., 0]
        m = gc.non_edge_mask(ref.shape, g)
        scale = np.abs(ref[m]).max()
        # the minimum of the squared distances is exact; the distance itself is one fused
        # multiply-add sequence and a square root away from the reference's: 4 ulp
        assert np.abs(mine - ref)[m].max() <= 1e-15 * scale * 4
    # a run with the computed distance reproduces the reference's history (SST reads it)
    gpu.close()


def test_wall_distance_is_a_noop_without_points():
    import aither_b200
    prob = refcase.problem_from_dump(gc.load("viscousFlatPlate"))
    gpu = aither_b200.GridLevel(prob)
    before = gpu.field(0, abi.FIELD_WALL_DIST).copy()
    gpu.compute_wall_distance(np.zeros((0, 3)))
    assert np.array_equal(before, gpu.field(0, abi.FIELD_WALL_DIST))
    gpu.close()
