"""Grid metrics of a structured block from its Plot3D nodes, in the reference's layout.

Set-up code (runs once, on the host, in numpy): cell volumes, face-area vectors, centroids, face
centres and cell widths, with the ghost-cell geometry mirrored from the interior. Follows reference
src/plot3d.cpp:60-360 (Volume, Centroid, FaceArea{I,J,K}, FaceCenter{I,J,K}),
src/procBlock.cpp:2160-2240 (AssignGhostCellsGeom) and :6397-6411 (CalcCellWidths).

Arrays are returned ghost padded, (k, j, i, component) with i fastest, i.e. exactly what
include/aither_gpu.h's aither_block_desc expects. Edge ghost cells (ghost in two directions) are
filled by mirroring as well; they are only read by viscous/gradient stencils.
"""
import numpy as np


def _pyramid_volume(p, a, b, c, d):
    # reference src/plot3d.cpp PyramidVolume: 1/6 * (p - a/b/c/d average) . (diag1 x diag2)
    xp = 0.25 * ((a - p) + (b - p) + (c - p) + (d - p))
    xac = c - a
    xbd = d - b
    return (1.0 / 6.0) * np.einsum("...i,...i->...", xp, np.cross(xac, xbd))


def interior_metrics(nodes):
    """nodes: (nk+1, nj+1, ni+1, 3). Returns dict of un-padded metric arrays."""
    x = np.asarray(nodes, dtype=np.float64)
    c000, c100 = x[:-1, :-1, :-1], x[:-1, :-1, 1:]
    c010, c110 = x[:-1, 1:, :-1], x[:-1, 1:, 1:]
    c001, c101 = x[1:, :-1, :-1], x[1:, :-1, 1:]
    c011, c111 = x[1:, 1:, :-1], x[1:, 1:, 1:]
    # centroid = mean of the 8 nodes (src/plot3d.cpp Centroid(ii,jj,kk))
    cen = 0.125 * (c000 + c100 + c010 + c110 + c001 + c101 + c011 + c111)
    # volume = sum of 6 pyramids (src/plot3d.cpp:60-112); node order as in the reference
    vol = (_pyramid_volume(cen, c000, c001, c011, c010) +
           _pyramid_volume(cen, c100, c110, c111, c101) +
           _pyramid_volume(cen, c000, c100, c101, c001) +
           _pyramid_volume(cen, c010, c011, c111, c110) +
           _pyramid_volume(cen, c000, c010, c110, c100) +
           _pyramid_volume(cen, c001, c101, c111, c011))

    def unit_mag(vec):
        mag = np.sqrt(np.einsum("...i,...i->...", vec, vec))
        return np.concatenate([vec / mag[..., None], mag[..., None]], axis=-1)

    # i-faces (src/plot3d.cpp:152-184): 0.5 * xbd x xac
    xi = x[:, :, :, :]
    xac = xi[1:, 1:, :, :] - xi[:-1, :-1, :, :]
    xbd = xi[:-1, 1:, :, :] - xi[1:, :-1, :, :]
    fAI = unit_mag(0.5 * np.cross(xbd, xac))
    fCI = 0.25 * (xi[:-1, :-1] + xi[:-1, 1:] + xi[1:, :-1] + xi[1:, 1:])
    # j-faces (:225-257)
    xac = x[1:, :, :-1, :] - x[:-1, :, 1:, :]
    xbd = x[:-1, :, :-1, :] - x[1:, :, 1:, :]
    fAJ = unit_mag(0.5 * np.cross(xbd, xac))
    fCJ = 0.25 * (x[:-1, :, :-1] + x[:-1, :, 1:] + x[1:, :, :-1] + x[1:, :, 1:])
    # k-faces (:300-332)
    xac = x[:, 1:, :-1, :] - x[:, :-1, 1:, :]
    xbd = x[:, 1:, 1:, :] - x[:, :-1, :-1, :]
    fAK = unit_mag(0.5 * np.cross(xbd, xac))
    fCK = 0.25 * (x[:, :-1, :-1] + x[:, :-1, 1:] + x[:, 1:, :-1] + x[:, 1:, 1:])
    return dict(vol=vol, center=cen, fAreaI=fAI, fAreaJ=fAJ, fAreaK=fAK, fCenterI=fCI,
                fCenterJ=fCJ, fCenterK=fCK)


def _pad(a, g, extra=(0, 0, 0)):
    pad = [(g, g), (g, g), (g, g)] + [(0, 0)] * (a.ndim - 3)
    return np.pad(a, pad, mode="constant")


def block_metrics(nodes, g, interblock_faces=()):
    """Ghost-padded metrics for one block.

    `interblock_faces`: surface types (1..6) whose ghost geometry comes from a neighbour block and
    is therefore NOT mirrored here (the caller swaps it in, as the reference does with
    SwapGeomSlice). All other ghost layers are filled by reflection of the interior layers
    (AssignGhostCellsGeom: volume and face areas copied from the `layer`-th interior cell, centres
    shifted by the interior spacing).
    """
    m = interior_metrics(nodes)
    nk, nj, ni = m["vol"].shape
    out = {k: _pad(v, g) for k, v in m.items()}
    dims = {"i": ni, "j": nj, "k": nk}
    axis = {"i": 2, "j": 1, "k": 0}

    def take(arr, ax, idx):
        sl = [slice(None)] * arr.ndim
        sl[ax] = idx
        return arr[tuple(sl)]

    def put(arr, ax, idx, val):
        sl = [slice(None)] * arr.ndim
        sl[ax] = idx
        arr[tuple(sl)] = val

    # regular ghosts first in i, then j, then k; later directions also extend over the already
    # filled ghost strips, which produces the edge-ghost geometry by double reflection.
    for d in ("i", "j", "k"):
        ax = axis[d]
        n = dims[d]
        fa_own = {"i": "fAreaI", "j": "fAreaJ", "k": "fAreaK"}[d]
        fc_own = {"i": "fCenterI", "j": "fCenterJ", "k": "fCenterK"}[d]
        for upper in (False, True):
            surf = {"i": 1, "j": 3, "k": 5}[d] + (1 if upper else 0)
            if surf in interblock_faces:
                continue
            for layer in range(1, g + 1):
                if upper:
                    gc, ic, pc = n + layer - 1, max(n - layer, 0), n + layer - 2
                    pic = ic + 1
                    iface, gface = max(n - layer, 0), n + layer
                    piface = iface + 1
                else:
                    gc, ic, pc = -layer, min(layer - 1, n - 1), -layer + 1
                    pic = ic - 1
                    iface, gface = min(layer, n), -layer
                    piface = iface - 1
                G = lambda q: q + g  # physical index -> padded index
                # volumes, cell-centred face areas of the other two directions
                put(out["vol"], ax, G(gc), take(out["vol"], ax, G(ic)))
                for name in ("fAreaI", "fAreaJ", "fAreaK"):
                    if name == fa_own:
                        # own-direction faces: ghost face `gface` mirrors interior face `iface`
                        put(out[name], ax, G(gface), take(out[name], ax, G(iface)))
                    else:
                        put(out[name], ax, G(gc), take(out[name], ax, G(ic)))
                distF2F = take(out[fc_own], ax, G(piface)) - take(out[fc_own], ax, G(iface))
                if layer > 1:
                    distC2C = take(out["center"], ax, G(pic)) - take(out["center"], ax, G(ic))
                else:
                    distC2C = None
                # own-direction face centres move by the face-to-face distance
                pface = gface - 1 if upper else gface + 1
                put(out[fc_own], ax, G(gface), take(out[fc_own], ax, G(pface)) + distF2F)
                # centroids
                if distC2C is None:
                    # first layer: use the face distance (reference uses distF2F)
                    shift_c = distF2F
                else:
                    shift_c = distC2C
                put(out["center"], ax, G(gc), take(out["center"], ax, G(pc)) +
                    _match(shift_c, take(out["center"], ax, G(pc))))
                for name in ("fCenterI", "fCenterJ", "fCenterK"):
                    if name == fc_own:
                        continue
                    prev = take(out[name], ax, G(pc))
                    put(out[name], ax, G(gc), prev + _grow(shift_c, prev))
    # cell widths from face centres (CalcCellWidths), every padded cell
    def width(fc, ax):
        lo = take(fc, ax, slice(0, fc.shape[ax] - 1))
        hi = take(fc, ax, slice(1, fc.shape[ax]))
        return np.sqrt(((hi - lo) ** 2).sum(axis=-1))

    out["cellWidthI"] = width(out["fCenterI"], 2)
    out["cellWidthJ"] = width(out["fCenterJ"], 1)
    out["cellWidthK"] = width(out["fCenterK"], 0)
    return out


def _match(shift, target):
    return _grow(shift, target)


def _grow(shift, target):
    """Extend `shift` (a 2-D slab) by repeating its last row/column so it matches `target`
    (face arrays are one longer in their own direction; reference GrowI/J/K)."""
    s = shift
    for ax in range(2):
        if s.shape[ax] < target.shape[ax]:
            last = np.take(s, [-1], axis=ax)
            s = np.concatenate([s, last], axis=ax)
        elif s.shape[ax] > target.shape[ax]:
            s = np.take(s, range(target.shape[ax]), axis=ax)
    return s


def pad_face_arrays(m, g):
    """The padded face arrays must be one longer in their own direction: already so, since the
    interior face arrays are. Provided for symmetry / documentation."""
    return m
