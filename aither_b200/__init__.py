"""aither_b200 -- B200-native hot path of the Aither structured-grid flow solver.

The product is the C-ABI shared library `aither_b200/lib/libaither_b200.so` (CUDA, sm_100a; ABI in
include/aither_gpu.h). This package is the thin Python host side used by the tests and bench:
`GridLevel` mirrors the reference's `gridLevel` / `mgSolution::Iterate` interface (reference
include/gridLevel.hpp:84-109, src/mgSolution.cpp:246-269) one method per phase, calling straight
through ctypes. There is no CPU fallback: if the library or a GPU is missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import ctypes_abi as abi
from .problem import Block, Problem, make_cfg  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
# AITHER_B200_LIB: another build of the same library (bisect builds of scripts/build_bisect.sh)
LIB_PATH = os.environ.get("AITHER_B200_LIB") or os.path.join(_HERE, "lib", "libaither_b200.so")
_LIB = None

# every symbol include/aither_gpu.h declares
ABI_SYMBOLS = (
    "aither_gpu_create", "aither_gpu_store_old_solution", "aither_gpu_iterate",
    "aither_gpu_get_boundary_conditions", "aither_gpu_calc_residual", "aither_gpu_calc_time_step",
    "aither_gpu_invert_diagonal", "aither_gpu_initialize_matrix_update", "aither_gpu_relax",
    "aither_gpu_update_blocks", "aither_gpu_reset_diagonal", "aither_gpu_run",
    "aither_gpu_upload_state", "aither_gpu_upload_state_async", "aither_gpu_upload_state_commit",
    "aither_gpu_upload_interior_async",
    "aither_gpu_download_state", "aither_gpu_download_field", "aither_gpu_download_wall_data",
    "aither_gpu_download_output", "aither_gpu_compute_wall_distance",
    "aither_gpu_field_size", "aither_gpu_synchronize", "aither_gpu_timer_start",
    "aither_gpu_timer_stop", "aither_gpu_launch_count", "aither_gpu_profile_enable",
    "aither_gpu_profile_get", "aither_gpu_kernel_family_name", "aither_gpu_num_kernel_families",
    "aither_gpu_destroy", "aither_gpu_last_error", "aither_gpu_version",
    "aither_gpu_alloc_host", "aither_gpu_free_host",
    "aither_gpu_comm_unique_id", "aither_gpu_comm_create", "aither_gpu_comm_destroy",
    "aither_gpu_halo_info", "aither_gpu_halo_p2p_export", "aither_gpu_halo_p2p_import",
    "aither_gpu_set_transfer", "aither_gpu_mg_restrict", "aither_gpu_mg_save_update",
    "aither_gpu_mg_subtract_saved", "aither_gpu_mg_prolong",
)


class AitherGpuError(RuntimeError):
    pass


def load_library():
    """dlopen the CUDA library and declare its prototypes. Raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise AitherGpuError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
                             % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    pd = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.aither_gpu_create.argtypes = [C.POINTER(abi.Cfg), C.c_int, C.POINTER(abi.BlockDesc), C.c_int,
                                    C.POINTER(abi.Conn), C.c_int, C.c_int, vp, C.c_int,
                                    C.POINTER(vp)]
    L.aither_gpu_set_transfer.argtypes = [vp, C.c_int, C.POINTER(C.c_int), pd, pd]
    L.aither_gpu_mg_restrict.argtypes = [vp, vp, C.c_int, C.c_double]
    L.aither_gpu_mg_save_update.argtypes = [vp]
    L.aither_gpu_mg_subtract_saved.argtypes = [vp]
    L.aither_gpu_mg_prolong.argtypes = [vp, vp]
    L.aither_gpu_store_old_solution.argtypes = [vp, C.c_int]
    L.aither_gpu_iterate.argtypes = [vp, C.c_double, C.c_int, pd, C.POINTER(abi.Linf), pd]
    for name in ("aither_gpu_get_boundary_conditions", "aither_gpu_calc_residual",
                 "aither_gpu_invert_diagonal", "aither_gpu_initialize_matrix_update",
                 "aither_gpu_reset_diagonal", "aither_gpu_synchronize", "aither_gpu_timer_start",
                 "aither_gpu_destroy"):
        getattr(L, name).argtypes = [vp]
    L.aither_gpu_calc_time_step.argtypes = [vp, C.c_double]
    L.aither_gpu_relax.argtypes = [vp, C.c_int, pd]
    L.aither_gpu_update_blocks.argtypes = [vp, C.c_int, pd, C.POINTER(abi.Linf)]
    L.aither_gpu_run.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, pd]
    L.aither_gpu_upload_state.argtypes = [vp, C.c_int, pd]
    L.aither_gpu_upload_state_async.argtypes = [vp, C.c_int, pd]
    L.aither_gpu_upload_state_commit.argtypes = [vp]
    L.aither_gpu_upload_interior_async.argtypes = [vp, C.c_int, pd]
    L.aither_gpu_download_state.argtypes = [vp, C.c_int, pd]
    L.aither_gpu_download_field.argtypes = [vp, C.c_int, C.c_int, pd]
    L.aither_gpu_download_wall_data.argtypes = [vp, C.c_int, C.c_int, pd]
    L.aither_gpu_download_output.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, pd]
    L.aither_gpu_compute_wall_distance.argtypes = [vp, pd, C.c_longlong]
    L.aither_gpu_halo_p2p_export.argtypes = [vp, C.c_char_p]
    L.aither_gpu_halo_p2p_import.argtypes = [vp, C.c_char_p]
    L.aither_gpu_field_size.argtypes = [vp, C.c_int, C.c_int]
    L.aither_gpu_field_size.restype = C.c_longlong
    L.aither_gpu_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.aither_gpu_launch_count.argtypes = [vp]
    L.aither_gpu_launch_count.restype = C.c_longlong
    L.aither_gpu_profile_enable.argtypes = [vp, C.c_int]
    L.aither_gpu_profile_get.argtypes = [vp, C.c_int, pd, C.POINTER(C.c_longlong)]
    L.aither_gpu_kernel_family_name.argtypes = [C.c_int]
    L.aither_gpu_kernel_family_name.restype = C.c_char_p
    L.aither_gpu_last_error.restype = C.c_char_p
    L.aither_gpu_alloc_host.argtypes = [C.c_longlong, C.POINTER(vp)]
    L.aither_gpu_free_host.argtypes = [vp]
    L.aither_gpu_version.restype = C.c_char_p
    L.aither_gpu_comm_unique_id.argtypes = [C.c_char_p]
    L.aither_gpu_comm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.aither_gpu_comm_destroy.argtypes = [vp]
    L.aither_gpu_halo_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
    _LIB = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pinned_array(shape):
    """float64 numpy array in page-locked host memory (aither_gpu_alloc_host); keeps itself alive
    through the array's base object and is freed with the library when the process exits."""
    L = load_library()
    n = int(np.prod(shape))
    p = C.c_void_p()
    if L.aither_gpu_alloc_host(n * 8, C.byref(p)) != 0:
        raise AitherGpuError(L.aither_gpu_last_error().decode())
    buf = (C.c_double * n).from_address(p.value)
    return np.ctypeslib.as_array(buf).reshape(shape)


class GridLevel:
    """Device-resident grid level + linear solver for the blocks owned by this rank."""

    def __init__(self, problem, device=0, rank=0, n_ranks=1, block_ids=None, nccl_comm=None):
        self.problem = problem
        self.neq = problem.neq
        self.block_ids = list(range(len(problem.blocks))) if block_ids is None else list(block_ids)
        self._lib = load_library()
        descs, conns, keep = problem.c_records(self.block_ids)
        h = C.c_void_p()
        rc = self._lib.aither_gpu_create(C.byref(problem.cfg), len(self.block_ids), descs,
                                         len(problem.conns), conns, rank, n_ranks, nccl_comm,
                                         device, C.byref(h))
        self._h = h if rc == 0 else None
        self._check(rc)

    def _check(self, rc):
        if rc != 0:
            raise AitherGpuError(self._lib.aither_gpu_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.aither_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mgSolution / gridLevel interface --------------------------------------------------
    def store_old_solution(self, it=0):
        self._check(self._lib.aither_gpu_store_old_solution(self._h, it))

    def get_boundary_conditions(self):
        self._check(self._lib.aither_gpu_get_boundary_conditions(self._h))

    def calc_residual(self):
        self._check(self._lib.aither_gpu_calc_residual(self._h))

    def calc_time_step(self, cfl):
        self._check(self._lib.aither_gpu_calc_time_step(self._h, cfl))

    def invert_diagonal(self):
        self._check(self._lib.aither_gpu_invert_diagonal(self._h))

    def initialize_matrix_update(self):
        self._check(self._lib.aither_gpu_initialize_matrix_update(self._h))

    def relax(self, sweeps=None):
        if sweeps is None:
            sweeps = self.problem.cfg.matrixSweeps
        mr = C.c_double()
        self._check(self._lib.aither_gpu_relax(self._h, sweeps, C.byref(mr)))
        return mr.value

    def update_blocks(self, mm=0):
        l2 = np.zeros(self.neq)
        linf = abi.Linf()
        self._check(self._lib.aither_gpu_update_blocks(self._h, mm, _ptr(l2), C.byref(linf)))
        return l2, linf

    def reset_diagonal(self):
        self._check(self._lib.aither_gpu_reset_diagonal(self._h))

    def iterate(self, cfl, mm=0):
        """One nonlinear iteration; returns (residL2[neq], linf, matrixResid)."""
        l2 = np.zeros(self.neq)
        linf = abi.Linf()
        mr = C.c_double()
        self._check(self._lib.aither_gpu_iterate(self._h, cfl, mm, _ptr(l2), C.byref(linf),
                                                 C.byref(mr)))
        return l2, linf, mr.value

    def run(self, n_iter, cfl_start, cfl_step=0.0, cfl_max=None):
        """n_iter time steps back to back; returns hist[n_iter, neq + 1]."""
        hist = np.zeros((n_iter * max(1, self.problem.cfg.nonlinearIterations), self.neq + 1))
        self._check(self._lib.aither_gpu_run(self._h, n_iter, cfl_start, cfl_step,
                                             cfl_start if cfl_max is None else cfl_max, _ptr(hist)))
        return hist

    # ---- data movement ---------------------------------------------------------------------
    def field(self, blk, fld):
        n = self._lib.aither_gpu_field_size(self._h, blk, fld)
        if n < 0:
            raise AitherGpuError(self._lib.aither_gpu_last_error().decode())
        out = np.empty(n)
        self._check(self._lib.aither_gpu_download_field(self._h, blk, fld, _ptr(out)))
        b = self.problem.blocks[self.block_ids[blk]]
        g = self.problem.cfg.numGhosts
        padded = fld in (abi.FIELD_STATE, abi.FIELD_UPDATE, abi.FIELD_TEMPERATURE,
                         abi.FIELD_VISCOSITY, abi.FIELD_EDDY_VISCOSITY, abi.FIELD_F1,
                         abi.FIELD_F2, abi.FIELD_VELOCITY_GRAD, abi.FIELD_WALL_DIST)
        shp = b.padded_shape(g) if padded else (b.nk, b.nj, b.ni)
        return out.reshape(shp + (-1,))

    def output(self, blk, var, scale=1.0, species=0):
        """one function-file variable (abi.OUT_*) of block `blk`, derived on the device, physical
        cells only: shape (nk, nj, ni)"""
        b = self.problem.blocks[self.block_ids[blk]]
        out = np.empty((b.nk, b.nj, b.ni))
        self._check(self._lib.aither_gpu_download_output(self._h, blk, var, species, scale,
                                                         _ptr(out)))
        return out

    def enable_peer_exchange(self):
        """ghost exchange over NVLink peer memory instead of NCCL send / recv (ranks of one node;
        collective: every rank calls it). AITHER_B200_HALO_P2P=0 keeps NCCL."""
        import os
        from . import distributed as adist
        if os.environ.get("AITHER_B200_HALO_P2P", "1") == "0":
            return False
        mine = C.create_string_buffer(64)
        self._check(self._lib.aither_gpu_halo_p2p_export(self._h, mine))
        everyone = adist.all_gather_bytes(mine.raw, 64)
        self._check(self._lib.aither_gpu_halo_p2p_import(self._h, everyone))
        return True

    def compute_wall_distance(self, wall_face_centers):
        """wall distance of every block from the centres (n, 3) of all viscous-wall faces, on the
        device (replaces the set-up's k-d tree search)"""
        pts = np.ascontiguousarray(wall_face_centers, dtype=np.float64).reshape(-1, 3)
        self._check(self._lib.aither_gpu_compute_wall_distance(self._h, _ptr(pts), pts.shape[0]))

    def wall_data(self, blk, surface):
        """wall variables (y+, shear stress, heat flux, T, mu_t, mu, rho, u_tau, k, omega) of a
        wall-law surface of block `blk`, shape (nk, nj, ni, 12) over the surface's cell range"""
        sf = self.problem.blocks[self.block_ids[blk]].surfaces[surface]
        shp = (max(sf[6] - sf[5], 1), max(sf[4] - sf[3], 1), max(sf[2] - sf[1], 1))
        out = np.empty(shp + (12,))
        self._check(self._lib.aither_gpu_download_wall_data(self._h, blk, surface, _ptr(out)))
        return out

    def upload_state(self, blk, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        self._check(self._lib.aither_gpu_upload_state(self._h, blk, _ptr(state)))

    def upload_state_async(self, blk, state):
        """start the host-to-device copy of `state` (page-locked, C-contiguous float64; it must
        stay alive and unchanged until `upload_state_commit`) on the copy stream"""
        assert state.flags["C_CONTIGUOUS"] and state.dtype == np.float64
        self._check(self._lib.aither_gpu_upload_state_async(self._h, blk, _ptr(state)))

    def upload_interior_async(self, blk, interior):
        """as upload_state_async for the physical cells only (nk, nj, ni, neq): the ghost cells are
        filled at the start of every iteration and need not cross PCIe"""
        if not interior.flags["C_CONTIGUOUS"] or interior.dtype != np.float64:
            raise ValueError("upload_interior_async needs a C-contiguous float64 array")
        self._check(self._lib.aither_gpu_upload_interior_async(self._h, blk, _ptr(interior)))

    def upload_state_commit(self):
        self._check(self._lib.aither_gpu_upload_state_commit(self._h))

    def download_state_into(self, blk, out):
        self._check(self._lib.aither_gpu_download_state(self._h, blk, _ptr(out)))

    # ---- timing ----------------------------------------------------------------------------
    def synchronize(self):
        self._check(self._lib.aither_gpu_synchronize(self._h))

    def timer_start(self):
        self._check(self._lib.aither_gpu_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self._lib.aither_gpu_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return self._lib.aither_gpu_launch_count(self._h)

    def profile_enable(self, on=True):
        self._check(self._lib.aither_gpu_profile_enable(self._h, int(on)))

    def profile(self):
        """{family: (ms, launches)} accumulated since profile_enable."""
        out = {}
        for f in range(self._lib.aither_gpu_num_kernel_families()):
            ms, n = C.c_double(), C.c_longlong()
            self._check(self._lib.aither_gpu_profile_get(self._h, f, C.byref(ms), C.byref(n)))
            out[self._lib.aither_gpu_kernel_family_name(f).decode()] = (ms.value, n.value)
        return out


class Multigrid:
    """Device-resident multigrid solution: one GridLevel per grid level (finest first) and the
    reference's full-approximation-storage cycle (mgSolution::Iterate / ImplicitUpdate /
    CycleAtLevel, src/mgSolution.cpp:160-269) composed from the per-level phase calls and the
    transfer operators of the C ABI (aither_gpu_mg_*).

    `problems`: one Problem per level (the coarse blocks are built once by the caller, as the
    reference's gridLevel::Coarsen does at set-up); `transfers[l]`: per block of level l the maps
    onto level l + 1 -- (toCoarse int32 [nk, nj, ni, 3], volWeightFactor [nk, nj, ni],
    prolongCoeffs [nk, nj, ni, 7]); `cycle_index`: 1 = V cycle, 2 = W cycle."""

    def __init__(self, problems, transfers, cycle_index, device=0):
        self.levels = [GridLevel(p, device=device) for p in problems]
        self.cycle_index = int(cycle_index)
        self.neq = problems[0].neq
        self._lib = load_library()
        self._keep = []
        for l, per_block in enumerate(transfers):
            for bb, (tc, vf, pc) in enumerate(per_block):
                tc = np.ascontiguousarray(tc, dtype=np.int32)
                vf = np.ascontiguousarray(vf, dtype=np.float64)
                pc = np.ascontiguousarray(pc, dtype=np.float64)
                lv = self.levels[l]
                lv._check(self._lib.aither_gpu_set_transfer(
                    lv._h, bb, tc.ctypes.data_as(C.POINTER(C.c_int)), _ptr(vf), _ptr(pc)))

    def close(self):
        for lv in self.levels:
            lv.close()

    def store_old_solution(self, it=0):
        self.levels[0].store_old_solution(it)

    def _cycle(self, fl, mm, cfl):
        lv = self.levels
        sweeps = lv[0].problem.cfg.matrixSweeps
        if fl == len(lv) - 1:
            return lv[fl].relax(sweeps)
        half = max(sweeps // 2, 1)
        lv[fl].relax(half)
        lv[fl]._check(self._lib.aither_gpu_mg_restrict(lv[fl]._h, lv[fl + 1]._h, mm, cfl))
        lv[fl + 1]._check(self._lib.aither_gpu_mg_save_update(lv[fl + 1]._h))
        for _ in range(self.cycle_index):
            self._cycle(fl + 1, mm, cfl)
        lv[fl + 1]._check(self._lib.aither_gpu_mg_subtract_saved(lv[fl + 1]._h))
        lv[fl]._check(self._lib.aither_gpu_mg_prolong(lv[fl + 1]._h, lv[fl]._h))
        return lv[fl].relax(half)

    def iterate(self, cfl, mm=0):
        f = self.levels[0]
        f.get_boundary_conditions()
        f.calc_residual()
        f.calc_time_step(cfl)
        f.invert_diagonal()
        f.initialize_matrix_update()
        mr = self._cycle(0, mm, cfl)
        l2, linf = f.update_blocks(mm)
        for lv in self.levels:
            lv.reset_diagonal()
        return l2, linf, mr
