"""Host-side description of one grid level: config + blocks + connections.

This is the Python stand-in for the reference objects a maintainer's shim would read
(`input`, `physics`, `gridLevel::Blocks()`, `gridLevel::Connections()`): it only carries arrays in
the reference's own layout (ghost-padded array-of-structs, i fastest; reference
include/multiArray3d.hpp:96-126) and turns them into the POD records of include/aither_gpu.h.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import ctypes_abi as abi

BLOCK_ARRAYS = ("state", "vol", "fAreaI", "fAreaJ", "fAreaK", "center", "cellWidthI", "cellWidthJ",
                "cellWidthK", "wallDist")


@dataclass
class Block:
    ni: int
    nj: int
    nk: int
    surfaces: List[Tuple[int, int, int, int, int, int, int, int]]  # type,imin,imax,jmin,jmax,kmin,kmax,tag
    arrays: Dict[str, np.ndarray]
    parent_block: int = 0
    global_pos: int = 0

    def padded_shape(self, g):
        return (self.nk + 2 * g, self.nj + 2 * g, self.ni + 2 * g)


@dataclass
class Problem:
    cfg: abi.Cfg
    blocks: List[Block]
    conns: List[abi.Conn] = field(default_factory=list)

    @property
    def neq(self):
        return self.cfg.neq

    @property
    def num_cells(self):
        return sum(b.ni * b.nj * b.nk for b in self.blocks)

    def c_records(self, block_ids: Sequence[int] = None):
        """(BlockDesc array, Conn array, keepalive) for the given (default: all) blocks."""
        ids = list(range(len(self.blocks))) if block_ids is None else list(block_ids)
        keep = []
        descs = (abi.BlockDesc * len(ids))()
        for n, bi in enumerate(ids):
            b = self.blocks[bi]
            d = descs[n]
            d.ni, d.nj, d.nk = b.ni, b.nj, b.nk
            d.parentBlock, d.globalPos = b.parent_block, b.global_pos
            surfs = (abi.Surface * len(b.surfaces))()
            for s, row in enumerate(b.surfaces):
                (surfs[s].type, surfs[s].imin, surfs[s].imax, surfs[s].jmin, surfs[s].jmax,
                 surfs[s].kmin, surfs[s].kmax, surfs[s].tag) = [int(v) for v in row]
            keep.append(surfs)
            d.numSurfaces = len(b.surfaces)
            d.surfaces = surfs
            for name in BLOCK_ARRAYS:
                arr = b.arrays.get(name)
                if arr is None:
                    setattr(d, name, None)
                    continue
                arr = np.ascontiguousarray(arr, dtype=np.float64)
                keep.append(arr)
                setattr(d, name, arr.ctypes.data_as(C.POINTER(C.c_double)))
        conns = (abi.Conn * max(1, len(self.conns)))()
        for n, c in enumerate(self.conns):
            conns[n] = c
        keep.append(conns)
        return descs, conns, keep


def make_cfg(**kw):
    """Build an `aither_cfg` from keyword arguments; list values fill the fixed-size arrays."""
    cfg = abi.Cfg()
    bc_states = kw.pop("bcStates", [])
    for k, v in kw.items():
        cur = getattr(cfg, k)
        if isinstance(cur, C.Array):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(cfg, k, v)
    cfg.numBCStates = len(bc_states)
    for i, st in enumerate(bc_states):
        rec = cfg.bcStates[i]
        for k, v in st.items():
            cur = getattr(rec, k)
            if isinstance(cur, C.Array):
                for j, x in enumerate(v):
                    cur[j] = x
            else:
                setattr(rec, k, v)
    return cfg
