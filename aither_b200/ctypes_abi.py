"""ctypes mirror of include/aither_gpu.h (POD structs and enums only, no library loading).

Kept separate from the library binding so that test infrastructure (the CPU oracle wrapper in
tests/) can build the same `aither_cfg` / `aither_block_desc` / `aither_conn` records that feed
the GPU path.
"""
import ctypes as C

MAX_SPECIES = 8
MAX_BC_STATES = 32

# enum aither_bc_type
BC_NONE, BC_SLIP_WALL, BC_VISCOUS_WALL, BC_CHARACTERISTIC, BC_INLET = 0, 1, 2, 3, 4
BC_SUPERSONIC_INFLOW, BC_SUPERSONIC_OUTFLOW, BC_STAGNATION_INLET = 5, 6, 7
BC_PRESSURE_OUTLET, BC_INTERBLOCK, BC_PERIODIC = 8, 9, 10
BC_NAMES = {
    "slipWall": BC_SLIP_WALL, "viscousWall": BC_VISCOUS_WALL,
    "characteristic": BC_CHARACTERISTIC, "inlet": BC_INLET,
    "supersonicInflow": BC_SUPERSONIC_INFLOW, "supersonicOutflow": BC_SUPERSONIC_OUTFLOW,
    "stagnationInlet": BC_STAGNATION_INLET, "pressureOutlet": BC_PRESSURE_OUTLET,
    "interblock": BC_INTERBLOCK, "periodic": BC_PERIODIC,
}

RECON_CONSTANT, RECON_MUSCL, RECON_WENO, RECON_WENOZ = 0, 1, 2, 3
LIMITER_NONE, LIMITER_VAN_ALBADA, LIMITER_MINMOD = 0, 1, 2
FLUX_ROE, FLUX_AUSM = 0, 1
JAC_RUSANOV, JAC_APPROX_ROE = 0, 1
SOLVER_LUSGS, SOLVER_DPLUR = 0, 1
TURB_NONE, TURB_KW_WILCOX, TURB_SST = 0, 1, 2

# enum aither_field
FIELD_STATE, FIELD_RESIDUAL, FIELD_SPEC_RADIUS, FIELD_DT, FIELD_DIAG = 0, 1, 2, 3, 4
FIELD_DIAG_INV, FIELD_UPDATE, FIELD_CONS_N, FIELD_MATRIX_RESID = 5, 6, 7, 8
FIELD_TEMPERATURE, FIELD_CONS_NM1, FIELD_VISCOSITY = 9, 10, 11
FIELD_EDDY_VISCOSITY, FIELD_F1, FIELD_F2, FIELD_VELOCITY_GRAD = 12, 13, 14, 15
FIELD_TKE_GRAD, FIELD_OMEGA_GRAD, FIELD_PRESSURE_GRAD = 16, 17, 18
FIELD_WALL_DIST = 19

_dS = C.c_double * MAX_SPECIES
_d3 = C.c_double * 3


class BCState(C.Structure):
    _fields_ = [
        ("tag", C.c_int), ("type", C.c_int),
        ("density", C.c_double), ("velocity", _d3), ("pressure", C.c_double),
        ("massFractions", _dS),
        ("stagnationPressure", C.c_double), ("stagnationTemperature", C.c_double),
        ("direction", _d3),
        ("temperature", C.c_double), ("heatFlux", C.c_double),
        ("isIsothermal", C.c_int), ("isConstantHeatFlux", C.c_int),
        ("turbulenceIntensity", C.c_double), ("eddyViscosityRatio", C.c_double),
        ("isWallLaw", C.c_int), ("vonKarmen", C.c_double), ("wallConstant", C.c_double),
        ("isNonreflecting", C.c_int), ("lengthScale", C.c_double),
    ]


class Cfg(C.Structure):
    _fields_ = [
        ("numSpecies", C.c_int), ("numTurb", C.c_int), ("numGhosts", C.c_int),
        ("isViscous", C.c_int), ("isRANS", C.c_int), ("isBlockMatrix", C.c_int),
        ("isMultilevelTime", C.c_int),
        ("recon", C.c_int), ("limiter", C.c_int), ("invFlux", C.c_int), ("invFluxJac", C.c_int),
        ("viscRecon", C.c_int), ("turbModel", C.c_int), ("solver", C.c_int),
        ("matrixSweeps", C.c_int), ("matrixRequiresInit", C.c_int),
        ("nonlinearIterations", C.c_int),
        ("kappa", C.c_double), ("theta", C.c_double), ("zeta", C.c_double),
        ("matrixRelaxation", C.c_double), ("dualTimeCFL", C.c_double),
        ("dtNondim", C.c_double), ("viscousCFLCoeff", C.c_double),
        ("gasConstant", _dS), ("n", _dS), ("hf", _dS),
        ("nondimScaling", C.c_double),
        ("suthViscC1", _dS), ("suthViscS", _dS), ("suthCondC1", _dS), ("suthCondS", _dS),
        ("molarMass", _dS),
        ("tRef", C.c_double), ("muMixRef", C.c_double), ("kMixRef", C.c_double),
        ("schmidt", C.c_double), ("turbPrandtl", C.c_double),
        ("numBCStates", C.c_int),
        ("bcStates", BCState * MAX_BC_STATES),
    ]

    @property
    def neq(self):
        return self.numSpecies + 4 + self.numTurb


class Surface(C.Structure):
    _fields_ = [("type", C.c_int), ("imin", C.c_int), ("imax", C.c_int), ("jmin", C.c_int),
                ("jmax", C.c_int), ("kmin", C.c_int), ("kmax", C.c_int), ("tag", C.c_int)]


_pd = C.POINTER(C.c_double)


class BlockDesc(C.Structure):
    _fields_ = [
        ("ni", C.c_int), ("nj", C.c_int), ("nk", C.c_int),
        ("parentBlock", C.c_int), ("globalPos", C.c_int),
        ("numSurfaces", C.c_int), ("surfaces", C.POINTER(Surface)),
        ("state", _pd), ("vol", _pd), ("fAreaI", _pd), ("fAreaJ", _pd), ("fAreaK", _pd),
        ("center", _pd), ("cellWidthI", _pd), ("cellWidthJ", _pd), ("cellWidthK", _pd),
        ("wallDist", _pd),
    ]


_i2 = C.c_int * 2


class Conn(C.Structure):
    _fields_ = [
        ("rank", _i2), ("block", _i2), ("localBlock", _i2), ("boundary", _i2),
        ("d1Start", _i2), ("d1End", _i2), ("d2Start", _i2), ("d2End", _i2),
        ("constSurf", _i2), ("patchBorder", C.c_int * 8),
        ("orientation", C.c_int), ("isInterblock", C.c_int),
    ]


class Linf(C.Structure):
    _fields_ = [("linf", C.c_double), ("block", C.c_int), ("i", C.c_int), ("j", C.c_int),
                ("k", C.c_int), ("eqn", C.c_int)]

# aither_output_var (include/aither_gpu.h)
(OUT_DENSITY, OUT_VEL_X, OUT_VEL_Y, OUT_VEL_Z, OUT_PRESSURE, OUT_MACH, OUT_SOS, OUT_DT,
 OUT_TEMPERATURE, OUT_ENERGY, OUT_ENTHALPY, OUT_CP, OUT_CV, OUT_VISCOSITY_RATIO,
 OUT_TURBULENT_VISCOSITY, OUT_VISCOSITY, OUT_TKE, OUT_SDR, OUT_F1, OUT_F2, OUT_WALL_DISTANCE,
 OUT_MASS_FRACTION) = range(22)
