"""Where does subsonicCylinder (BASELINE configs[0]) exceed the 1e-12 per-cell residual bar?

Bisect by inputs: (1) ghost cells after the BC phase against the reference's dump, per boundary
type; (2) the residual as the iteration computes it; (3) the residual with the REFERENCE's ghost
cells uploaded (state@it0.bc), i.e. the residual kernel alone on identical inputs. The CPU oracle
(reference formulas operation for operation, another compiler / libm) is run beside it.
Test infrastructure (uses tests/ and oracle/); run on the GPU box: python scripts/diag_subsonic.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldencheck as gc  # noqa: E402
import oracle  # noqa: E402
import refcase  # noqa: E402
import aither_b200  # noqa: E402
from aither_b200 import ctypes_abi as abi  # noqa: E402


NOISE = ()


def per_eq(a, b):
    ax = tuple(range(a.ndim - 1))
    scale = np.abs(b).max(axis=ax)
    scale = np.where(scale > 0, scale, 1.0)
    e = np.abs(a - b).max(axis=ax) / scale
    e[list(NOISE)] = 0.0  # out-of-plane momentum: rounding noise (goldencheck.noise_equations)
    return e


for name in sys.argv[1:] or ["subsonicCylinder"]:
    d = gc.load(name)
    NOISE = gc.noise_equations(d)
    prob = refcase.problem_from_dump(d, state_key="state0")
    g = prob.cfg.numGhosts
    out = {}
    for label, make in (("gpu", aither_b200.GridLevel), ("oracle", oracle.OracleLevel)):
        lvl = make(prob)
        lvl.store_old_solution(0)
        lvl.get_boundary_conditions()
        st = lvl.field(0, abi.FIELD_STATE)
        ref_bc = d["b0/state@it0.bc"].reshape(st.shape)
        m = (gc.non_corner_mask if prob.cfg.isViscous else gc.non_edge_mask)(st.shape[:3], g)
        err = np.abs(st - ref_bc) / np.abs(ref_bc).reshape(-1, st.shape[-1]).max(axis=0)
        err[~m] = 0.0
        w = np.unravel_index(np.argmax(err), err.shape)
        print("%s %-6s ghosts after BC: max rel err %.3e at (k,j,i,eq)=%s" % (name, label, err.max(), w))
        lvl.calc_residual()
        r = lvl.field(0, abi.FIELD_RESIDUAL)
        rref = d["b0/residual@it0"].reshape(r.shape)
        e = per_eq(r, rref)
        rr = np.abs(r - rref) / np.abs(rref).reshape(-1, r.shape[-1]).max(axis=0)
        rr[..., list(NOISE)] = 0.0
        w = np.unravel_index(np.argmax(rr), rr.shape)
        print("%s %-6s residual, own BC ghosts:        per equation %s   worst cell (k,j,i)=%s" %
              (name, label, np.array2string(e, precision=2), w[:3]))
        lvl.close()
        # the residual kernel alone: the reference's ghost cells as input
        prob2 = refcase.problem_from_dump(d, state_key="state@it0.bc")
        lvl = make(prob2)
        lvl.store_old_solution(0)
        lvl.calc_residual()
        r = lvl.field(0, abi.FIELD_RESIDUAL)
        print("%s %-6s residual, reference's BC ghosts: per equation %s" %
              (name, label, np.array2string(per_eq(r, rref), precision=2)))
        lvl.close()
