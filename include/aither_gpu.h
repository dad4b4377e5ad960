/* aither_gpu.h -- C ABI of the B200 hot path (libaither_b200.so).
 *
 * Drop-in boundary for the reference solver's per-iteration hot path
 * (mnucci32/aither v0.10.0). The reference has no plugin/FFI seam; the seam is
 * the body of mgSolution::Iterate (reference src/mgSolution.cpp:246-269), which
 * calls gridLevel::{GetBoundaryConditions,CalcResidual,CalcTimeStep,
 * InvertDiagonal,InitializeMatrixUpdate,Relax,UpdateBlocks,ResetDiagonal}
 * (reference include/gridLevel.hpp:84-109) and the abstract linearSolver
 * (reference include/linearSolver.hpp:37-93). A maintainer replaces that body
 * with aither_gpu_iterate(); see INTEGRATION.md for the shim.
 *
 * Conventions
 *   - plain C, POD structs, host pointers, no torch / CUDA types;
 *   - every entry point returns 0 on success, non-zero on error, and never
 *     calls exit() (the reference does: e.g. src/matrix.cpp:83-86); the message
 *     is available from aither_gpu_last_error();
 *   - host arrays are in the reference's own layout: array-of-structs, i
 *     fastest, ghost padded (reference include/multiArray3d.hpp:96-126), i.e.
 *       index = blk * ((i+g) + (j+g)*NI + (k+g)*NI*NJ) + l ,  NI = ni + 2g;
 *     the library converts to its device layout on upload;
 *   - variable order inside a cell: primitive [rho_1..rho_ns,u,v,w,p,(k,w)],
 *     conserved/residual/update [rho_1..,rho u,rho v,rho w,rho E,(rho k,rho w)]
 *     (reference include/varArray.hpp:47-51);
 *   - one host thread drives one handle; one handle drives one GPU.
 */
#ifndef AITHER_GPU_H
#define AITHER_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define AITHER_MAX_SPECIES 8
#define AITHER_MAX_BC_STATES 32

/* boundary-condition kinds: the reference selects these by string compare
 * (src/ghostStates.cpp:62-689, include/boundaryConditions.hpp:291). */
enum aither_bc_type {
  AITHER_BC_NONE = 0,
  AITHER_BC_SLIP_WALL = 1,
  AITHER_BC_VISCOUS_WALL = 2,
  AITHER_BC_CHARACTERISTIC = 3,
  AITHER_BC_INLET = 4,
  AITHER_BC_SUPERSONIC_INFLOW = 5,
  AITHER_BC_SUPERSONIC_OUTFLOW = 6,
  AITHER_BC_STAGNATION_INLET = 7,
  AITHER_BC_PRESSURE_OUTLET = 8,
  AITHER_BC_INTERBLOCK = 9,
  AITHER_BC_PERIODIC = 10
};

enum aither_recon { AITHER_RECON_CONSTANT = 0, AITHER_RECON_MUSCL = 1,
                    AITHER_RECON_WENO = 2, AITHER_RECON_WENOZ = 3 };
enum aither_limiter { AITHER_LIMITER_NONE = 0, AITHER_LIMITER_VAN_ALBADA = 1,
                      AITHER_LIMITER_MINMOD = 2 };
enum aither_inv_flux { AITHER_FLUX_ROE = 0, AITHER_FLUX_AUSM = 1 };
enum aither_inv_jac { AITHER_JAC_RUSANOV = 0, AITHER_JAC_APPROX_ROE = 1 };
enum aither_solver { AITHER_SOLVER_LUSGS = 0, AITHER_SOLVER_DPLUR = 1 };
enum aither_turb { AITHER_TURB_NONE = 0, AITHER_TURB_KW_WILCOX = 1,
                   AITHER_TURB_SST = 2 };

/* One nondimensional boundary-state record: the fields of the reference's
 * inputState hierarchy (include/inputStates.hpp:45-109) that GetGhostState
 * reads, after input::NondimensionalizeStateData. */
typedef struct aither_bc_state {
  int tag;
  int type;                      /* aither_bc_type the record was declared for */
  double density;
  double velocity[3];
  double pressure;
  double massFractions[AITHER_MAX_SPECIES];
  double stagnationPressure;
  double stagnationTemperature;
  double direction[3];
  double temperature;            /* isothermal wall */
  double heatFlux;               /* constant heat-flux wall */
  int isIsothermal;
  int isConstantHeatFlux;
  double turbulenceIntensity;
  double eddyViscosityRatio;
  int isWallLaw;                 /* viscousWall(wallTreatment=wallLaw): include/inputStates.hpp:343-369 */
  double vonKarmen;              /* 0.41 unless given */
  double wallConstant;           /* 5.5 unless given */
  int isNonreflecting;           /* inlet / pressureOutlet(nonreflecting=true): LODI relaxation,
                                    src/ghostStates.cpp:435-466, :614-643 */
  double lengthScale;            /* its relaxation length */
} aither_bc_state;

/* POD snapshot of the reference's `input` + `physics` objects: only what the
 * hot path branches on (src/input.cpp:674-721,1110-1144) or evaluates. */
typedef struct aither_cfg {
  int numSpecies;                /* ns; neq = ns + 4 + numTurb (kernels exist for ns = 1, 2, 3) */
  int numTurb;                   /* 0, or 2 for RANS */
  int numGhosts;                 /* input::NumberGhostLayers */
  int isViscous;
  int isRANS;
  int isBlockMatrix;             /* blusgs / bdplur */
  int isMultilevelTime;          /* bdf2 */
  int recon;                     /* aither_recon */
  int limiter;                   /* aither_limiter */
  int invFlux;                   /* aither_inv_flux */
  int invFluxJac;                /* aither_inv_jac */
  int viscRecon;                 /* 0 central, 1 centralFourth */
  int turbModel;                 /* aither_turb */
  int solver;                    /* aither_solver */
  int matrixSweeps;
  int matrixRequiresInit;        /* input::MatrixRequiresInitialization */
  int nonlinearIterations;       /* per time step (>= 1); U^(n-1) <- U^n after the last one
                                    of a multilevel scheme (src/gridLevel.cpp:418-431) */
  double kappa;
  double theta;                  /* Beam-Warming theta (input.cpp:256-270) */
  double zeta;
  double matrixRelaxation;
  double dualTimeCFL;            /* <= 0: no dual time stepping */
  double dtNondim;               /* > 0: global time step dt*aRef/lRef; else local (CFL) */
  double viscousCFLCoeff;
  /* fluid: calorically perfect ideal gas, per species (nondimensional) */
  double gasConstant[AITHER_MAX_SPECIES];
  double n[AITHER_MAX_SPECIES];           /* cv = R n, cp = R (n+1) */
  double hf[AITHER_MAX_SPECIES];          /* heat of formation */
  /* Sutherland transport (src/transport.cpp), used when isViscous */
  double nondimScaling;                   /* mu_ref / (rho_ref a_ref l_ref) */
  double suthViscC1[AITHER_MAX_SPECIES], suthViscS[AITHER_MAX_SPECIES];
  double suthCondC1[AITHER_MAX_SPECIES], suthCondS[AITHER_MAX_SPECIES];
  double molarMass[AITHER_MAX_SPECIES];   /* Wilke's mixing rule (any consistent unit) */
  double tRef, muMixRef, kMixRef;
  double schmidt;                /* species diffusion (`diffusionModel: schmidt`): Schmidt number;
                                    <= 0 for `diffusionModel: none`. The turbulent Schmidt number
                                    is the reference's constant 0.7 (include/turbulence.hpp:71) */
  double turbPrandtl;            /* unused: Pr_t follows the turbulence model (0.9; k-omega 2006 8/9) */
  int numBCStates;
  aither_bc_state bcStates[AITHER_MAX_BC_STATES];
} aither_cfg;

/* One boundary surface of a block: a row of the reference's BC table
 * (`name imin imax jmin jmax kmin kmax tag`, node indices;
 * include/boundaryConditions.hpp:55-75). */
typedef struct aither_surface {
  int type;                      /* aither_bc_type */
  int imin, imax, jmin, jmax, kmin, kmax;
  int tag;
} aither_surface;

/* One procBlock (reference include/procBlock.hpp:64-124). All pointers are
 * host arrays owned by the caller, read during aither_gpu_create only. */
typedef struct aither_block_desc {
  int ni, nj, nk;                /* physical cells */
  int parentBlock;
  int globalPos;                 /* position in the decomposed block list */
  int numSurfaces;
  const aither_surface *surfaces;
  const double *state;           /* (nk+2g)(nj+2g)(ni+2g) x neq, primitive */
  const double *vol;             /* ghost padded, 1 */
  const double *fAreaI;          /* ghost padded, (ni+1) faces in i, 4 = {nx,ny,nz,|A|} */
  const double *fAreaJ;
  const double *fAreaK;
  const double *center;          /* ghost padded, 3 */
  const double *cellWidthI;      /* ghost padded, 1 */
  const double *cellWidthJ;
  const double *cellWidthK;
  const double *wallDist;        /* ghost padded, 1 (may be NULL when inviscid) */
} aither_block_desc;

/* The 12 fields of the reference's `connection`
 * (include/boundaryConditions.hpp:324-336), verbatim. Index [0] = first side,
 * [1] = second side. */
typedef struct aither_conn {
  int rank[2];
  int block[2];                  /* global block numbers */
  int localBlock[2];
  int boundary[2];               /* surface type 1..6 */
  int d1Start[2], d1End[2];
  int d2Start[2], d2End[2];
  int constSurf[2];
  int patchBorder[8];
  int orientation;               /* 1..8 */
  int isInterblock;              /* 0 = periodic */
} aither_conn;

/* L-infinity residual record: reference include/resid.hpp:24-31. */
typedef struct aither_linf {
  double linf;
  int block, i, j, k, eqn;       /* eqn is 1-based, as in procBlock.cpp:862-867 */
} aither_linf;

/* fields retrievable with aither_gpu_download_field (tests / output). Layout
 * of the returned host array is the reference's for that field. */
enum aither_field {
  AITHER_FIELD_STATE = 0,        /* ghost padded, neq */
  AITHER_FIELD_RESIDUAL = 1,     /* no ghosts, neq */
  AITHER_FIELD_SPEC_RADIUS = 2,  /* no ghosts, 2 = {flow, turb} */
  AITHER_FIELD_DT = 3,           /* no ghosts, 1 */
  AITHER_FIELD_DIAG = 4,         /* no ghosts, linearSolver::a_ block */
  AITHER_FIELD_DIAG_INV = 5,     /* no ghosts, linearSolver::aInv_ block */
  AITHER_FIELD_UPDATE = 6,       /* ghost padded, neq: linearSolver::x_ */
  AITHER_FIELD_CONS_N = 7,       /* no ghosts, neq */
  AITHER_FIELD_MATRIX_RESID = 8, /* no ghosts, neq: f - (Ax - b) */
  AITHER_FIELD_TEMPERATURE = 9,  /* ghost padded, 1 */
  AITHER_FIELD_CONS_NM1 = 10,    /* no ghosts, neq */
  AITHER_FIELD_VISCOSITY = 11,   /* ghost padded, 1 (viscous runs) */
  /* RANS runs: cell averages of the six face values (procBlock.cpp:1396-1452) */
  AITHER_FIELD_EDDY_VISCOSITY = 12, /* ghost padded, 1 */
  AITHER_FIELD_F1 = 13,          /* ghost padded, 1 */
  AITHER_FIELD_F2 = 14,          /* ghost padded, 1 */
  AITHER_FIELD_VELOCITY_GRAD = 15, /* ghost padded, 9: (r,c) = d u_c / d x_r */
  AITHER_FIELD_TKE_GRAD = 16,    /* no ghosts, 3 */
  AITHER_FIELD_OMEGA_GRAD = 17,  /* no ghosts, 3 */
  AITHER_FIELD_PRESSURE_GRAD = 18, /* no ghosts, 3: cell average of the face pressure gradients
                                     (kept for runs with non-reflecting BCs only) */
  AITHER_FIELD_WALL_DIST = 19    /* ghost padded, 1 (viscous runs) */
};

typedef struct aither_gpu aither_gpu;   /* opaque handle */

/* Build the device-side gridLevel (blocks + connections + linear solver
 * storage) for the blocks this rank owns. Replaces gridLevel's constructor and
 * input::AssignLinearSolver (reference src/gridLevel.cpp:50-120,
 * src/input.cpp:843-858). `ncclComm` is an ncclComm_t or NULL when every
 * connection is local. */
int aither_gpu_create(const aither_cfg *cfg, int nLocalBlocks,
                      const aither_block_desc *blocks, int nConnections,
                      const aither_conn *conns, int rank, int nRanks,
                      void *ncclComm, int device, aither_gpu **out);

/* mgSolution::StoreOldSolution (src/mgSolution.cpp:103-114): U^n <- state,
 * and U^{n-1} <- U^n on the first step of a multilevel scheme. */
int aither_gpu_store_old_solution(aither_gpu *h, int iter);

/* One nonlinear iteration: mgSolution::Iterate (src/mgSolution.cpp:246-269).
 * Outputs are what main.cpp:249-264 consumes: residL2[neq] = sum over local
 * cells of R^2 (un-rooted, procBlock.cpp:858), linf, and the matrix residual
 * sum(mr^2)/size exactly as mgSolution::CycleAtLevel returns it (:199-206). */
int aither_gpu_iterate(aither_gpu *h, double cfl, int mm, double *residL2,
                       aither_linf *linf, double *matrixResid);

/* The phases of aither_gpu_iterate, individually (gridLevel.hpp:84-109);
 * used by the parity tests to observe each phase boundary. */
int aither_gpu_get_boundary_conditions(aither_gpu *h);
int aither_gpu_calc_residual(aither_gpu *h);
int aither_gpu_calc_time_step(aither_gpu *h, double cfl);
int aither_gpu_invert_diagonal(aither_gpu *h);
int aither_gpu_initialize_matrix_update(aither_gpu *h);
int aither_gpu_relax(aither_gpu *h, int sweeps, double *matrixResid);
int aither_gpu_update_blocks(aither_gpu *h, int mm, double *residL2,
                             aither_linf *linf);
int aither_gpu_reset_diagonal(aither_gpu *h);

/* `nIter` time steps back to back without host synchronisation in between,
 * each of cfg.nonlinearIterations nonlinear iterations (residual norms of every
 * nonlinear iteration land in hist: (nIter * nonlinearIterations) x (neq + 1),
 * the last column being the matrix residual). Same arithmetic as aither_gpu_iterate;
 * the call main.cpp's loop would make when the log is only read at the end. */
int aither_gpu_run(aither_gpu *h, int nIter, double cflStart, double cflStep,
                   double cflMax, double *hist);

/* Output staging: one variable of the reference's function file (`outputVariables`, WriteFunFile
 * src/output.cpp:209-437) derived ON THE DEVICE for the physical cells of a block and copied to
 * `dst` ((nk, nj, ni) doubles, i fastest: the order WriteFunFile writes), so that only what is
 * asked for crosses PCIe instead of the ghost-padded state. `scale` is the dimensional factor the
 * reference multiplies by (e.g. rRef * aRef * aRef for pressure); the value is formed with the
 * reference's expression and then multiplied once. `species` selects the species for
 * AITHER_OUT_MASS_FRACTION. Variables whose field the configuration does not keep are refused. */
enum aither_output_var {
  AITHER_OUT_DENSITY = 0, AITHER_OUT_VEL_X, AITHER_OUT_VEL_Y, AITHER_OUT_VEL_Z, AITHER_OUT_PRESSURE,
  AITHER_OUT_MACH, AITHER_OUT_SOS, AITHER_OUT_DT, AITHER_OUT_TEMPERATURE, AITHER_OUT_ENERGY,
  AITHER_OUT_ENTHALPY, AITHER_OUT_CP, AITHER_OUT_CV, AITHER_OUT_VISCOSITY_RATIO,
  AITHER_OUT_TURBULENT_VISCOSITY, AITHER_OUT_VISCOSITY, AITHER_OUT_TKE, AITHER_OUT_SDR,
  AITHER_OUT_F1, AITHER_OUT_F2, AITHER_OUT_WALL_DISTANCE, AITHER_OUT_MASS_FRACTION,
  AITHER_OUT_NUM_VARS
};
int aither_gpu_download_output(aither_gpu *h, int blk, int var, int species, double scale,
                               double *dst);

/* Wall distance on the device. Replaces the k-d tree search of the set-up:
 * kdtree::NearestNeighbor (src/kdtree.cpp:123-225) called for every physical cell by
 * procBlock::CalcWallDistance (src/procBlock.cpp:6030-6107) with the tree main.cpp builds from
 * GetViscousFaceCenters (src/main.cpp:144,191-201; src/utility.cpp:310-368). `wallFaceCenters`
 * holds the n centres (x, y, z) of ALL viscous-wall faces of the simulation -- every rank passes
 * the same list, as the reference broadcasts it. For every block of the handle the distance of
 * each physical cell centre to the nearest of them replaces the wall distance given at create,
 * and the ghost cells (not the edge ghost cells, which keep what they held) follow the
 * reference's rule: minus the mirrored interior value across a viscous wall, the value of the
 * first interior cell elsewhere. Exhaustive search, tiled through shared memory: the minimum is
 * the tree's minimum. n = 0 leaves everything as it is (the reference skips the call). */
int aither_gpu_compute_wall_distance(aither_gpu *h, const double *wallFaceCenters, long long n);

/* Wall variables of one viscous-wall surface (reference `wallData` / `wallVars`,
 * include/wallData.hpp:40-57; what WriteWallFunFile reads, src/output.cpp:440-588): for every face
 * of surface `surface` (index into the block's surface list as given to aither_gpu_create) the
 * AITHER_WALL_VARS doubles {y+, shear stress x/y/z, heat flux, wall temperature, wall eddy
 * viscosity, wall viscosity, wall density, friction velocity, k, omega}, in the reference's face
 * order (i fastest, then j, then k over the surface's cell range), as the last residual
 * evaluation left them. Kept on the device for walls with `wallTreatment: wallLaw`; any other
 * surface is refused. */
#define AITHER_WALL_VARS 12
int aither_gpu_download_wall_data(aither_gpu *h, int blk, int surface, double *dst);

/* host <-> device state transfer in the reference's layout
 * (procBlock::States(), include/procBlock.hpp:506). */
int aither_gpu_upload_state(aither_gpu *h, int blk, const double *stateAoS);
int aither_gpu_download_state(aither_gpu *h, int blk, double *stateAoS);
/* Pipelined variant for a host that owns the state and hands a fresh copy over every step:
 * _async starts the host-to-device copy (from page-locked memory) on the library's copy stream
 * and returns at once, so the copy overlaps the iteration in flight; _commit makes the compute
 * stream wait for the copy and converts it into the device layout in front of the next
 * aither_gpu_iterate. One upload may be pending per handle. */
int aither_gpu_upload_state_async(aither_gpu *h, int blk, const double *stateAoS);
int aither_gpu_upload_state_commit(aither_gpu *h);
/* The same for a host that hands over the physical cells only (nk x nj x ni x neq, the
 * reference's layout without the ghost shell): the ghost cells are filled at the start of every
 * iteration (gridLevel::GetBoundaryConditions), so they need not cross PCIe. Committed with
 * aither_gpu_upload_state_commit. */
int aither_gpu_upload_interior_async(aither_gpu *h, int blk, const double *interiorAoS);
int aither_gpu_download_field(aither_gpu *h, int blk, int field, double *dst);
/* number of doubles aither_gpu_download_field writes for `field` */
long long aither_gpu_field_size(aither_gpu *h, int blk, int field);

/* page-locked host buffers for the state transfers above (cudaMallocHost /
 * cudaFreeHost): a caller that wants the copies at full PCIe rate allocates its
 * staging arrays here. Usable before any handle exists. */
int aither_gpu_alloc_host(long long bytes, void **out);
int aither_gpu_free_host(void *p);

/* device synchronisation + CUDA-event timing helpers for bench.py (the
 * launching stream is the library's own, which torch.cuda.Event cannot see). */
int aither_gpu_synchronize(aither_gpu *h);
int aither_gpu_timer_start(aither_gpu *h);
int aither_gpu_timer_stop(aither_gpu *h, float *milliseconds);
/* kernels launched by this handle since creation */
long long aither_gpu_launch_count(aither_gpu *h);
/* per-kernel-family accumulated device time (ms) and launch counts since the
 * last reset; families are listed by aither_gpu_kernel_family_name */
int aither_gpu_profile_enable(aither_gpu *h, int enable);
int aither_gpu_profile_get(aither_gpu *h, int family, double *ms, long long *launches);
const char *aither_gpu_kernel_family_name(int family);
int aither_gpu_num_kernel_families(void);

/* Multi-GPU: one process (reference "rank") per GPU. Connections whose two sides
 * live on different ranks exchange their ghost layers with ncclSend/ncclRecv
 * (replacing MPI_Sendrecv_replace, reference include/multiArray3d.hpp:1440-1508).
 * The host program makes the communicator the way it would an MPI one: rank 0
 * calls aither_gpu_comm_unique_id, broadcasts the 128 bytes with whatever it has
 * (MPI_Bcast in the reference's main.cpp, torch.distributed in bench.py), every
 * rank calls aither_gpu_comm_create and passes the result to aither_gpu_create.
 * libnccl.so.2 is bound with dlopen on first use; single-GPU runs never load it. */
int aither_gpu_comm_unique_id(char id[128]);
int aither_gpu_comm_create(const char id[128], int rank, int nRanks, int device,
                           void **comm);
int aither_gpu_comm_destroy(void *comm);
/* Direct ghost exchange over NVLink peer memory instead of ncclSend / ncclRecv (which cost ~95 us
 * per exchange whatever the size): the donor rank's pack kernel writes a slice straight into the
 * acceptor rank's receive buffer and raises a flag behind it; the acceptor's stream waits for the
 * flag and unpacks. Replaces the MPI_Sendrecv of multiArray3d::SwapSliceMPI
 * (include/multiArray3d.hpp:1440-1508) like the NCCL path, bit for bit the same ghost cells.
 * Every rank (one process per GPU of ONE node) calls _export after aither_gpu_create, the host
 * gathers the AITHER_P2P_HANDLE_BYTES-byte handles of all ranks in rank order with the transport
 * it already has (MPI_Allgather), and every rank calls _import with the gathered block. Without
 * these two calls the exchange stays on NCCL. */
#define AITHER_P2P_HANDLE_BYTES 64
int aither_gpu_halo_p2p_export(aither_gpu *h, void *handle);
int aither_gpu_halo_p2p_import(aither_gpu *h, const void *handles);
/* number of halo levels (pack/unpack launch pairs per exchange) and doubles this
 * rank sends to other ranks per component per exchange; for reports and tests */
int aither_gpu_halo_info(aither_gpu *h, int *levels, long long *remoteCells);

/* ---- multigrid transfer operators between two levels ----------------------
 * Every grid level is a handle of its own, created from the coarse blocks the reference builds
 * once (gridLevel::Coarsen, setup). These calls replace gridLevel::Restriction / Prolongation
 * (src/gridLevel.cpp:538-611) and linearSolver::SubtractFromUpdate (src/linearSolver.cpp:195-201);
 * the cycle (mgSolution::CycleAtLevel, src/mgSolution.cpp:160-207) is composed by the caller from
 * them and the per-level phase calls above, as the reference composes it from gridLevel methods.
 *   set_transfer : maps of one block of the FINE level onto the next coarser level, in the
 *                  reference's layouts -- toCoarse_ (nk nj ni 3 ints), volWeightFactor_ (nk nj ni),
 *                  prolongCoeffs_ (nk nj ni 7) (include/gridLevel.hpp:56-58)
 *   mg_restrict  : volume-weighted state -> coarse; coarse BCs, residual, time step, diagonal;
 *                  volume-weighted update -> coarse; forcing = (A x - b) + summed fine matrix
 *                  residual (of the fine level's last aither_gpu_relax), folded into the coarse
 *                  right-hand side
 *   mg_save_update / mg_subtract_saved : coarseDu = x before the coarse cycles, x -= coarseDu after
 *   mg_prolong   : node-averaged trilinear interpolation of the coarse update, added to the fine */
int aither_gpu_set_transfer(aither_gpu *fine, int blk, const int *toCoarse, const double *volFac,
                            const double *prolongCoeffs);
int aither_gpu_mg_restrict(aither_gpu *fine, aither_gpu *coarse, int mm, double cfl);
int aither_gpu_mg_save_update(aither_gpu *h);
int aither_gpu_mg_subtract_saved(aither_gpu *h);
int aither_gpu_mg_prolong(aither_gpu *coarse, aither_gpu *fine);

int aither_gpu_destroy(aither_gpu *h);
const char *aither_gpu_last_error(void);
const char *aither_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AITHER_GPU_H */
