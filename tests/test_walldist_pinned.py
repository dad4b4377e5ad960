"""The numpy restatement of the reference's wall-distance set-up (tests/walldist_ref.py) against
the reference's own wallDist_ arrays (tests/golden/walldist_*.npz, generator make_walldist.py)."""
import numpy as np
import pytest

import goldencheck as gc
import walldist_ref as wr


def viscous_wall_id():
    from aither_b200 import ctypes_abi as abi
    return abi.BC_VISCOUS_WALL


@pytest.mark.parametrize("name", ["viscousFlatPlate", "wallLaw", "couette"])
def test_numpy_wall_distance_matches_reference(name):
    d = wr.load(name)
    wall = viscous_wall_id()
    pts = wr.wall_face_centers(d, wall)
    assert len(pts) > 0
    for bb in range(int(d["numBlocks"][0])):
        ref = d["b%d/wallDist" % bb][..., 0]
        g = int(d["b%d/dims" % bb][3])
        mine, defined = wr.wall_distance(d, bb, pts, wall)
        m = gc.non_edge_mask(ref.shape, g) & defined
        scale = np.abs(ref[m]).max()
        assert np.abs(mine - ref)[m].max() <= 1e-14 * scale
