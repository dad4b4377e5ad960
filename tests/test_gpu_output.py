"""Output staging (-m gpu): aither_gpu_download_output derives the function-file variables of the
reference (WriteFunFile, src/output.cpp:229-330) on the device for the physical cells only. Checked
against the same expressions evaluated in numpy from the downloaded state (and against the device
fields the reference writes verbatim: temperature, viscosity, eddy viscosity, F1, F2, dt)."""
import numpy as np
import pytest

from aither_b200 import ctypes_abi as abi
from aither_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_output_variables_euler():
    import aither_b200
    prob = synthetic.box_problem(20, 12, 9, seed=31, amplitude=0.03)
    g = prob.cfg.numGhosts
    gpu = aither_b200.GridLevel(prob)
    for it in range(2):
        gpu.store_old_solution(it)
        gpu.iterate(40.0)
    st = gpu.field(0, abi.FIELD_STATE)[g:-g, g:-g, g:-g]
    rho, u, v, w, p = (st[..., q] for q in range(5))
    R, n = prob.cfg.gasConstant[0], prob.cfg.n[0]
    gamma = (R * (n + 1.0)) / (R * n)
    sos = np.sqrt(gamma * p / rho)
    t = p / (rho * R)
    vel2 = u * u + v * v + w * w
    scale = 3.7  # the dimensional factor is applied once, after the reference's expression
    want = {
        abi.OUT_DENSITY: rho, abi.OUT_VEL_X: u, abi.OUT_VEL_Y: v, abi.OUT_VEL_Z: w,
        abi.OUT_PRESSURE: p, abi.OUT_MACH: np.sqrt(vel2) / sos, abi.OUT_SOS: sos,
        abi.OUT_TEMPERATURE: t, abi.OUT_ENERGY: R * n * t + 0.5 * vel2,
        abi.OUT_ENTHALPY: R * (n + 1.0) * t + 0.5 * vel2,
        abi.OUT_CP: np.full_like(rho, R * (n + 1.0)), abi.OUT_CV: np.full_like(rho, R * n),
        abi.OUT_MASS_FRACTION: np.ones_like(rho),
        abi.OUT_DT: gpu.field(0, abi.FIELD_DT)[..., 0],
    }
    for var, ref in want.items():
        mine = gpu.output(0, var, scale=scale)
        assert mine.shape == ref.shape
        assert np.abs(mine - scale * ref).max() <= 1e-14 * np.abs(scale * ref).max(), var
    for var in (abi.OUT_VISCOSITY, abi.OUT_F1, abi.OUT_WALL_DISTANCE):
        with pytest.raises(aither_b200.AitherGpuError):
            gpu.output(0, var)   # not kept by an Euler run: refused, never zero-filled
    gpu.close()


def test_output_variables_rans():
    import aither_b200
    prob = synthetic.box_problem(16, 12, 8, seed=33, amplitude=0.02, turb="sst2003",
                                 limiter="vanAlbada", size=16 * 1e-4)
    g = prob.cfg.numGhosts
    gpu = aither_b200.GridLevel(prob)
    gpu.store_old_solution(0)
    gpu.iterate(20.0)
    gpu.get_boundary_conditions()
    gpu.calc_residual()
    cut = lambda a: a[g:-g, g:-g, g:-g, 0]
    mu, mut = cut(gpu.field(0, abi.FIELD_VISCOSITY)), cut(gpu.field(0, abi.FIELD_EDDY_VISCOSITY))
    st = gpu.field(0, abi.FIELD_STATE)[g:-g, g:-g, g:-g]
    want = {abi.OUT_VISCOSITY: mu, abi.OUT_TURBULENT_VISCOSITY: mut,
            abi.OUT_VISCOSITY_RATIO: mut / mu, abi.OUT_F1: cut(gpu.field(0, abi.FIELD_F1)),
            abi.OUT_F2: cut(gpu.field(0, abi.FIELD_F2)), abi.OUT_TKE: st[..., 5],
            abi.OUT_SDR: st[..., 6]}
    for var, ref in want.items():
        mine = gpu.output(0, var, scale=2.0)
        assert np.abs(mine - 2.0 * ref).max() <= 1e-14 * max(np.abs(2.0 * ref).max(), 1e-300), var
    gpu.close()
