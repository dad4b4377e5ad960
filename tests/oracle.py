"""ctypes wrapper of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from aither_b200 import ctypes_abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        src = os.path.join(ROOT, "oracle", "aither_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.Cfg), C.c_int, C.POINTER(abi.BlockDesc), C.c_int,
                                 C.POINTER(abi.Conn)]
        L.orc_destroy.argtypes = [C.c_void_p]
        for name in ("orc_get_boundary_conditions", "orc_calc_residual", "orc_invert_diagonal",
                     "orc_initialize_matrix_update", "orc_reset_diagonal"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.orc_store_old_solution.argtypes = [C.c_void_p, C.c_int]
        L.orc_calc_time_step.argtypes = [C.c_void_p, C.c_double]
        L.orc_relax.argtypes = [C.c_void_p, C.c_int]
        L.orc_relax.restype = C.c_double
        L.orc_update_blocks.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double),
                                        C.POINTER(abi.Linf)]
        L.orc_iterate.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double),
                                  C.POINTER(abi.Linf)]
        L.orc_iterate.restype = C.c_double
        L.orc_field_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_field_size.restype = C.c_longlong
        L.orc_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_get_wall_data.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_get_wall_data.restype = C.c_longlong
        L.orc_set_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        pd = C.POINTER(C.c_double)
        L.orc_muscl.argtypes = [pd, pd, pd, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double,
                                C.c_double, pd]
        L.orc_weno.argtypes = [C.POINTER(pd), pd, C.c_int, C.c_int, pd]
        L.orc_inviscid_flux.argtypes = [C.POINTER(abi.Cfg), pd, pd, pd, pd]
        L.orc_ghost_state.argtypes = [C.POINTER(abi.Cfg), pd, C.c_int, pd, C.c_int, C.c_int,
                                      C.c_int, pd]
        L.orc_offdiag_scalar.argtypes = [C.POINTER(abi.Cfg), pd, pd, pd, C.c_int, pd]
        L.orc_set_transfer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), pd, pd]
        L.orc_mg_restrict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.orc_mg_prolong.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_mg_save_update.argtypes = [C.c_void_p]
        L.orc_mg_subtract_saved.argtypes = [C.c_void_p]
        for name in ("orc_set_transfer", "orc_mg_restrict", "orc_mg_prolong", "orc_mg_save_update",
                     "orc_mg_subtract_saved"):
            getattr(L, name).restype = None
        _LIB = L
    return _LIB


class OracleLevel:
    """The oracle's gridLevel: same phase methods as the GPU binding (aither_b200.GridLevel)."""

    def __init__(self, problem):
        self.problem = problem
        self.neq = problem.neq
        descs, conns, keep = problem.c_records()
        self._keep = keep
        self._h = lib().orc_create(C.byref(problem.cfg), len(problem.blocks), descs,
                                   len(problem.conns), conns)

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def store_old_solution(self, it=0):
        lib().orc_store_old_solution(self._h, it)

    def get_boundary_conditions(self):
        lib().orc_get_boundary_conditions(self._h)

    def calc_residual(self):
        lib().orc_calc_residual(self._h)

    def calc_time_step(self, cfl):
        lib().orc_calc_time_step(self._h, cfl)

    def invert_diagonal(self):
        lib().orc_invert_diagonal(self._h)

    def initialize_matrix_update(self):
        lib().orc_initialize_matrix_update(self._h)

    def relax(self, sweeps=None):
        if sweeps is None:
            sweeps = self.problem.cfg.matrixSweeps
        return lib().orc_relax(self._h, sweeps)

    def update_blocks(self, mm=0):
        l2 = np.zeros(self.neq)
        linf = abi.Linf()
        lib().orc_update_blocks(self._h, mm, l2.ctypes.data_as(C.POINTER(C.c_double)),
                                C.byref(linf))
        return l2, linf

    def reset_diagonal(self):
        lib().orc_reset_diagonal(self._h)

    def iterate(self, cfl, mm=0):
        l2 = np.zeros(self.neq)
        linf = abi.Linf()
        mr = lib().orc_iterate(self._h, cfl, mm, l2.ctypes.data_as(C.POINTER(C.c_double)),
                               C.byref(linf))
        return l2, linf, mr

    def wall_data(self, blk, surface):
        """wall variables of a viscous-wall surface, shape (nk, nj, ni, 12) over its cell range"""
        sf = self.problem.blocks[blk].surfaces[surface]
        shp = (max(sf[6] - sf[5], 1), max(sf[4] - sf[3], 1), max(sf[2] - sf[1], 1))
        out = np.zeros(shp + (12,))
        n = lib().orc_get_wall_data(self._h, blk, surface, out.ctypes.data_as(C.POINTER(C.c_double)))
        assert n in (0, int(np.prod(shp))), (n, shp)
        return out if n else None

    def field(self, blk, fld):
        n = lib().orc_field_size(self._h, blk, fld)
        out = np.empty(n)
        lib().orc_get_field(self._h, blk, fld, out.ctypes.data_as(C.POINTER(C.c_double)))
        b = self.problem.blocks[blk]
        g = self.problem.cfg.numGhosts
        padded = fld in (abi.FIELD_STATE, abi.FIELD_UPDATE, abi.FIELD_TEMPERATURE,
                         abi.FIELD_VISCOSITY, abi.FIELD_EDDY_VISCOSITY, abi.FIELD_F1,
                         abi.FIELD_F2, abi.FIELD_VELOCITY_GRAD)
        shp = b.padded_shape(g) if padded else (b.nk, b.nj, b.ni)
        return out.reshape(shp + (-1,))


class OracleMultigrid:
    """The oracle's mgSolution: one OracleLevel per grid level (finest first) plus the transfer
    maps of every level pair, and the full-approximation-storage cycle of the reference
    (mgSolution::Iterate / ImplicitUpdate / CycleAtLevel, src/mgSolution.cpp:160-269) composed
    from the per-level phases and the oracle's transfer operators.

    `problems`: one Problem per level; `transfers[l]`: per block of level l a tuple
    (toCoarse int32 [nk, nj, ni, 3], volWeightFactor [nk, nj, ni], prolongCoeffs [nk, nj, ni, 7])
    onto level l + 1; `cycle_index`: 1 = V cycle, 2 = W cycle."""

    def __init__(self, problems, transfers, cycle_index):
        self.levels = [OracleLevel(p) for p in problems]
        self.cycle_index = cycle_index
        self.neq = problems[0].neq
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
        for l, per_block in enumerate(transfers):
            for bb, (tc, vf, pc) in enumerate(per_block):
                tc = np.ascontiguousarray(tc, dtype=np.int32)
                vf = np.ascontiguousarray(vf, dtype=np.float64)
                pc = np.ascontiguousarray(pc, dtype=np.float64)
                lib().orc_set_transfer(self.levels[l]._h, bb, tc.ctypes.data_as(pi),
                                       vf.ctypes.data_as(pd), pc.ctypes.data_as(pd))

    def close(self):
        for l in self.levels:
            l.close()

    def store_old_solution(self, it=0):
        self.levels[0].store_old_solution(it)

    def _cycle(self, fl, mm, cfl):
        lv = self.levels
        sweeps = lv[0].problem.cfg.matrixSweeps
        if fl == len(lv) - 1:
            return lv[fl].relax(sweeps)
        half = max(sweeps // 2, 1)
        lv[fl].relax(half)
        lib().orc_mg_restrict(lv[fl]._h, lv[fl + 1]._h, mm, cfl)
        lib().orc_mg_save_update(lv[fl + 1]._h)
        for _ in range(self.cycle_index):
            self._cycle(fl + 1, mm, cfl)
        lib().orc_mg_subtract_saved(lv[fl + 1]._h)
        lib().orc_mg_prolong(lv[fl + 1]._h, lv[fl]._h)
        return lv[fl].relax(half)

    def iterate(self, cfl, mm=0):
        f = self.levels[0]
        f.get_boundary_conditions()
        f.calc_residual()
        lib().orc_calc_time_step(f._h, cfl)
        lib().orc_invert_diagonal(f._h)
        lib().orc_initialize_matrix_update(f._h)
        mr = self._cycle(0, mm, cfl)
        l2, linf = f.update_blocks(mm)
        for l in self.levels:
            l.reset_diagonal()
        return l2, linf, mr
