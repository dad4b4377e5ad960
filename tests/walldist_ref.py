"""Wall distance the way the reference's set-up defines it, in numpy (TEST INFRASTRUCTURE).

ref: GetViscousFaceCenters (src/utility.cpp:310-368), kdtree::NearestNeighbor
(src/kdtree.cpp:123-225: the smallest squared distance, returned as its square root),
procBlock::CalcWallDistance (src/procBlock.cpp:6030-6107: physical cells by the search, ghost
cells -- not the edge ghost cells -- minus the mirrored interior value across a viscous wall, the
first interior cell's value elsewhere).
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VISCOUS_WALL = 1  # ref_harness.cpp BcTypeId / include/aither_gpu.h AITHER_BC_VISCOUS_WALL


def load(name):
    return np.load(os.path.join(GOLDEN, "walldist_%s.npz" % name))


def wall_face_centers(d, viscous_wall_id):
    """face centres of every viscous-wall face of every block, in the reference's order"""
    pts = []
    for bb in range(int(d["numBlocks"][0])):
        p = "b%d/" % bb
        g = int(d[p + "dims"][3])
        for typ, imin, imax, jmin, jmax, kmin, kmax, tag, stype in d[p + "surfaces"]:
            if typ != viscous_wall_id:
                continue
            if stype <= 2:
                fc = d[p + "fCenterI"]
                for jj in range(jmin, jmax):
                    for kk in range(kmin, kmax):
                        pts.append(fc[kk + g, jj + g, imin + g])
            elif stype <= 4:
                fc = d[p + "fCenterJ"]
                for ii in range(imin, imax):
                    for kk in range(kmin, kmax):
                        pts.append(fc[kk + g, jmin + g, ii + g])
            else:
                fc = d[p + "fCenterK"]
                for ii in range(imin, imax):
                    for jj in range(jmin, jmax):
                        pts.append(fc[kmin + g, jj + g, ii + g])
    return np.array(pts).reshape(-1, 3)


def wall_distance(d, bb, pts, viscous_wall_id, start=None, connection_ids=(9, 10)):
    """ghost-padded (K, J, I) wall distance of block bb and the mask of the cells it defines:
    physical cells and the ghost cells of boundary surfaces. Ghost cells across connections
    (swapped from the neighbour block afterwards, src/gridLevel.cpp:261-281) and edge ghost cells
    keep `start`'s values."""
    p = "b%d/" % bb
    ni, nj, nk, g = [int(v) for v in d[p + "dims"][:4]]
    c = d[p + "center"][g:g + nk, g:g + nj, g:g + ni]
    best = np.full(c.shape[:3], np.inf)
    for q in range(0, len(pts), 256):
        diff = c[..., None, :] - pts[q:q + 256]
        best = np.minimum(best, (diff * diff).sum(-1).min(-1))
    wd = np.zeros((nk + 2 * g, nj + 2 * g, ni + 2 * g)) if start is None else np.array(start, float)
    wd[g:g + nk, g:g + nj, g:g + ni] = np.sqrt(best)
    defined = np.zeros(wd.shape, bool)
    defined[g:g + nk, g:g + nj, g:g + ni] = True
    n = (ni, nj, nk)
    for typ, imin, imax, jmin, jmax, kmin, kmax, tag, stype in d[p + "surfaces"]:
        if typ in connection_ids:
            continue
        d3 = (stype - 1) // 2
        lo, hi = [imin, jmin, kmin], [imax, jmax, kmax]
        rng = [range(lo[q], hi[q]) for q in range(3)]
        wall = typ == viscous_wall_id
        for layer in range(1, g + 1):
            ghost = -layer if stype % 2 == 1 else n[d3] + layer - 1
            if stype % 2 == 1:
                src = layer - 1 if wall else 0
            else:
                src = n[d3] - layer if wall else n[d3] - 1
            rg, rs = list(rng), list(rng)
            rg[d3], rs[d3] = [ghost], [src]
            gi = np.ix_([k + g for k in rg[2]], [j + g for j in rg[1]], [i + g for i in rg[0]])
            si = np.ix_([k + g for k in rs[2]], [j + g for j in rs[1]], [i + g for i in rs[0]])
            wd[gi] = -wd[si] if wall else wd[si]
            defined[gi] = True
    return wd, defined
