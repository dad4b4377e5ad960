// aither_gpu.cu -- host side of the B200 hot path and its C ABI (include/aither_gpu.h).
//
// This file is the device-side `gridLevel` + `linearSolver`: it owns the blocks' HBM storage,
// sequences the kernels of one nonlinear iteration exactly as mgSolution::Iterate /
// ImplicitUpdate / CycleAtLevel do for a single grid level (ref: src/mgSolution.cpp:160-269), and
// returns what main.cpp's loop consumes. There is no CPU fallback: every entry point fails with an
// error if CUDA is unavailable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/aither_gpu.h"
#include "halo.cuh"
#include "kernels.cuh"
#include "march.cuh"
#include "implicit_tma.cuh"
#include "lusgs_wave.cuh"
#include "lusgs_pencil.cuh"
#include "viscous.cuh"

using namespace aither;

struct aither_gpu;
struct AitherEqOps {
  int (*PhaseBoundaryConditionsT)(aither_gpu *);
  int (*PhaseResidualT)(aither_gpu *, int, double);
  int (*PhasePrepT)(aither_gpu *, double, int);
  int (*PhaseRelaxT)(aither_gpu *, int, int);
  int (*PhaseUpdateT)(aither_gpu *, int, int);
  int (*StoreOldT)(aither_gpu *, int);
  int (*InitAuxT)(aither_gpu *, int);
  int (*OutputVarT)(aither_gpu *, int, int, int, double, double *);
};
const AitherEqOps *AitherEqOps_1_0();
const AitherEqOps *AitherEqOps_1_2();
const AitherEqOps *AitherEqOps_2_0();
const AitherEqOps *AitherEqOps_2_2();
const AitherEqOps *AitherEqOps_3_0();
const AitherEqOps *AitherEqOps_3_2();

// one last-error string per thread for the whole library (the equation-set translation units
// below share it: inline function, one instance after linking)
inline std::string &AitherLastError() {
  static thread_local std::string e;
  return e;
}
#define g_lastError AitherLastError()

// cuStreamWaitValue32 through the runtime's driver entry point (no link-time libcuda)
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
inline StreamWaitValue32Fn StreamWaitValueFn() {
  static StreamWaitValue32Fn fn = nullptr;
  static bool tried = false;
  if (tried) return fn;
  tried = true;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    fn = reinterpret_cast<StreamWaitValue32Fn>(p);
  return fn;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember it per device and
// kernel (one slot array per call site / template instantiation), under a lock -- handles on
// different GPUs may be created and driven from different host threads
#include <mutex>
inline std::mutex &AitherAttrMutex() {
  static std::mutex m;
  return m;
}
template <typename K>
int EnsureSmemOptIn(K kern, size_t bytes, int device, bool (&done)[64]) {
  std::lock_guard<std::mutex> lock(AitherAttrMutex());
  if (device < 0 || device >= 64) return 1;
  if (done[device]) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(bytes)) != cudaSuccess)
    return 1;
  done[device] = true;
  return 0;
}

namespace aither_host {  // host-side types shared by every translation unit of the library
enum Family {
  kFamBc = 0, kFamResidual, kFamPrep, kFamDplur, kFamLusgs, kFamAxmb, kFamUpdate, kFamStore,
  kFamReduce, kFamHalo, kFamLayout, kFamViscGhost, kFamViscFlux, kFamLusgsPack, kFamLusgsAhead, kNumFamilies
};
struct HostBlock {
  BlockDev dev;
  void *alloc = nullptr;          // one allocation holding every field
  size_t allocBytes = 0;
  std::vector<aither_surface> surfaces;
  std::vector<long long> surfFaceOffset;  // per surface: first boundary-face record, -1 for connections
  SurfDev *dSurfs = nullptr;
  int nBcSurfs = 0;
  EdgeSurf *dEdgeSurfs = nullptr;  // every surface of the block, connections included
  double *dWallVars = nullptr;     // wall-law records (walllaw.cuh), null without wall-law walls
  double *dPatchMach = nullptr;    // {average, maximum} Mach number per BC surface (non-reflecting BCs)
  // multigrid (multigrid.cuh): transfer maps onto the next coarser level (kept by the fine block),
  // child list per coarse cell (built on first use), saved update, node scratch of this block
  std::vector<int> toCoarseHost;
  int *dToCoarse = nullptr, *dChildren = nullptr;
  long long childrenFor = 0;       // coarse cell count the child list was built for
  double *dVolFac = nullptr, *dProlong = nullptr, *dSavedX = nullptr, *dNodes = nullptr;
  double *dSavedDiag = nullptr;    // diagonal of the previous restriction of this iteration
  int nEdgeSurfs = 0;
  long long bcThreads = 0;
  uint8_t *dConnFace[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int globalPos = 0;
  long long paddedCells = 0;
  dim3 cellGrid, cellBlock;       // cell-parallel kernels (32 x 8 threads)
  dim3 cell128Grid;               // register-heavy cell-parallel kernels (32 x 4 threads)
  dim3 updGrid;                   // update kernel: kUpdPlanes k-planes per block
  int nUpdBlocks = 0;
  dim3 resGrid, resBlock;
  int nCellBlocks = 0;
  // plane-marching kernels (march.cuh)
  dim3 marchGrid;
  int kChunk = 1;
  int nMarchBlocks = 0;
  // TMA-fed implicit sweep (implicit_tma.cuh)
  ImplMaps tmaMaps;
  dim3 tmaGrid;
  int tmaChunk = 1;
  int nTmaBlocks = 0;
  // every tile id of the TMA sweep, the tiles next to a block face with a connection first
  int *dTmaTiles = nullptr;
  int nTmaBoundaryTiles = 0;
  int nFields = 0;
  // LU-SGS: the plane launches of one half sweep, captured once as a CUDA graph
  // [forward / backward][first sweep form / full Gauss-Seidel]
  cudaGraphExec_t lusgsGraph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  // LU-SGS as a persistent wavefront (lusgs_wave.cuh): pencil order (anti-diagonals of the
  // (j, k) pencil lattice, sweep space) and the ticket + progress counters, built on first use
  int2 *dWaveOrder = nullptr;
  WaveSync *dWaveSync = nullptr;
  int waveTJ = 0, waveTK = 0, waveNbJ = 0, wavePencils = 0;
  size_t waveSyncBytes = 0;
  // scalar LU-SGS (lusgs_pencil.cuh): array-of-structs workspaces, one ghost layer included --
  // behind-side face areas for the forward / backward sweep (built once) and the per-iteration
  // record of what does not depend on the update
  double *dWaveGeoLo = nullptr, *dWaveGeoHi = nullptr, *dWaveDyn = nullptr, *dWaveAhead = nullptr;
  double *wallDistSlot = nullptr;
  bool ghostsInAlt = false;  // fused update: the ghost cells of the last fill sit in dev.stateAlt
  uint4 *dWaveMailJ = nullptr, *dWaveMailK = nullptr;  // hand-over between pencils (lusgs_pencil.cuh)
  unsigned waveTag = 0;                                 // number of the half sweep
  bool waveCarries = false;  // a half sweep leaves the next one's ahead-sums behind ...
  bool waveFixup = false;    // ... except next to connected faces (ghost cells exchanged in between)
};

}  // namespace aither_host
using namespace aither_host;

namespace {

const char *kFamilyNames[kNumFamilies] = {"bc_ghost_fill", "residual", "dt_diag_init", "dplur_sweep",
                                          "lusgs_plane", "matrix_residual", "update_norms",
                                          "store_time_n", "reduce_finalize", "halo_pack_unpack",
                                          "layout_convert", "viscous_ghosts_aux", "viscous_flux",
                                          "lusgs_pack", "lusgs_ahead"};

int Fail(const std::string &msg) {
  g_lastError = msg;
  return 1;
}

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      return Fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                  ":" + std::to_string(__LINE__) + ")");                               \
    }                                                                                  \
  } while (0)

}  // namespace

struct aither_gpu {
  aither_cfg cfg;
  Params params;
  int neq = 0, ns = 0, nt = 0, asz = 1;
  int device = 0;
  int rank = 0, nRanks = 1;
  cudaStream_t stream = nullptr;
  std::vector<HostBlock> blocks;
  std::vector<aither_conn> conns;
  HaloPlan halo;      // every connection with its tangential extension (the reference's swap)
  // direct exchange over peer memory: this rank's arena (receive buffers + flags) and the peers'
  unsigned char *p2pArena = nullptr;
  size_t p2pBytes = 0;
  std::vector<unsigned char *> p2pPeers;  // IPC mappings (nullptr at this rank's own index)
  HaloPlan haloFace;  // ghost cells straight behind the patches only: one level
  // ... and only the FIRST ghost layer: all that the implicit off-diagonals read of the update
  // (nearest neighbours; ref src/procBlock.cpp:1056-1170). Half the bytes of haloFace for MUSCL,
  // a third for WENO, on sweeps + 1 of the sweeps + 2 exchanges of an iteration.
  HaloPlan haloUpdate;
  bool updateOneLayer = true;  // AITHER_B200_HALO_LAYERS=all: exchange every ghost layer of the update (A/B)
  bool stateNeedsEdges = false;  // viscous stencils read the edge ghost cells of the state
  aither_bc_state *dBcStates = nullptr;
  double *dPartials = nullptr;    // per-thread-block partial sums
  LinfCand *dLinfPartials = nullptr;
  size_t partialsCap = 0;
  IterResult *dResults = nullptr;
  IterResult *hResults = nullptr;  // pinned
  int resultsCap = 0;
  double *dStage = nullptr;        // layout-conversion staging
  size_t stageBytes = 0;
  // pipelined state upload (aither_gpu_upload_state_async / _commit): its own staging buffer,
  // a copy stream, and the two events that order copy -> conversion -> next copy
  cudaStream_t copyStream = nullptr;
  double *dStage2 = nullptr;
  size_t stage2Bytes = 0;
  cudaEvent_t evCopied = nullptr, evConverted = nullptr;
  int pendingBlk = -1;
  bool pendingInterior = false;
  long long launches = 0;
  bool keepMatrixResid = false;
  // aither_gpu_iterate / _run: the matrix-residual pass also advances the state into the
  // alternate buffer (fuseUpdate: requested for this iteration; stateFused: done by PhaseRelax)
  bool fuseUpdate = false, stateFused = false;
  int jac = kJacScalar;            // JacKind of the implicit matrix
  bool consNStale = false;         // U^n not materialised (Params::timeTermsVanish)
  bool nonreflecting = false;      // some inlet / pressure outlet is non-reflecting
  bool wallLaw = false;            // some viscous wall uses the wall law
  bool lusgsGraphs = true;         // AITHER_B200_LUSGS_GRAPH=0: plain launches (A/B)
  bool lusgsSplit = true;          // AITHER_B200_LUSGS_SPLIT=0: one thread per cell (A/B)
  bool lusgsWave = true;           // AITHER_B200_LUSGS=planes: one launch per hyperplane (A/B)
  // ghost exchanges run on their own high-priority stream, ordered against the compute stream with
  // events; AITHER_B200_HALO_OVERLAP=1: the exchange of a DPLUR sweep's update runs beside the
  // sweep's interior tiles (A/B; default: every exchange in program order of the compute stream)
  cudaStream_t commStream = nullptr;
  cudaEvent_t evCompute = nullptr, evExchanged = nullptr;
  bool haloOverlap = true;
  // boundary tiles of the sweeps in flight bump this counter; the communication stream waits for
  // it with a stream memory operation (cuStreamWaitValue32: no SM is held while waiting)
  unsigned int *dHaloSignal = nullptr;
  unsigned int haloTarget = 0;
  bool stateMovedSinceStore = false;
  int *dFlag = nullptr;            // set by PrepBlockKernel on a singular diagonal block
  bool legacyKernels = false;      // AITHER_B200_KERNELS=legacy: the first-generation kernels
  bool tmaImplicit = true;         // AITHER_B200_KERNELS=march: register-fed implicit sweep
  bool fusePrep = true;            // AITHER_B200_FUSE_PREP=0: separate PrepKernel (A/B runs)
  // timing
  cudaEvent_t evStart = nullptr, evStop = nullptr;
  bool profile = false;
  struct Ev { cudaEvent_t a, b; int fam; };
  std::vector<Ev> evs;
  double famMs[kNumFamilies] = {0};
  long long famLaunches[kNumFamilies] = {0};
};

namespace {

struct ScopedLaunch {
  aither_gpu *h;
  int fam;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  ScopedLaunch(aither_gpu *h_, int fam_, cudaStream_t st_ = nullptr)
      : h(h_), fam(fam_), st(st_ ? st_ : h_->stream) {
    h->launches++;
    h->famLaunches[fam]++;
    if (h->profile) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~ScopedLaunch() {
    if (h->profile) {
      cudaEventRecord(b, st);
      h->evs.push_back({a, b, fam});
    }
  }
};

void DrainProfile(aither_gpu *h) {
  for (auto &e : h->evs) {
    float ms = 0.f;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&ms, e.a, e.b);
    h->famMs[e.fam] += ms;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  h->evs.clear();
}

int EnsureStage(aither_gpu *h, size_t bytes) {
  if (bytes <= h->stageBytes) return 0;
  if (h->dStage) cudaFree(h->dStage);
  if (h->dStage2) cudaFree(h->dStage2);
  if (h->evCopied) cudaEventDestroy(h->evCopied);
  if (h->evConverted) cudaEventDestroy(h->evConverted);
  if (h->copyStream) cudaStreamDestroy(h->copyStream);
  h->dStage = nullptr;
  h->stageBytes = 0;
  CK(cudaMalloc(&h->dStage, bytes));
  h->stageBytes = bytes;
  return 0;
}

// host AoS (reference layout, extents SI,SJ,SK with origin (oi,oj,ok) in physical index space)
// -> device SoA field
int UploadAos(aither_gpu *h, const HostBlock &hb, const double *src, int SI, int SJ, int SK,
              int nc, double *dstField, int originI, int originJ, int originK) {
  const size_t n = static_cast<size_t>(SI) * SJ * SK * nc;
  if (EnsureStage(h, n * sizeof(double))) return 1;
  CK(cudaMemcpyAsync(h->dStage, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const BlockDev &b = hb.dev;
  const long long cells = static_cast<long long>(SI) * SJ * SK;
  const int grid = static_cast<int>(std::min<long long>((cells + 255) / 256, 148 * 16));
  {
    ScopedLaunch sl(h, kFamLayout);
    AosToSoaKernel<<<grid, 256, 0, h->stream>>>(h->dStage, SI, SJ, SK, nc, dstField, b.fs,
                                                originI + b.lp, originJ + b.g, originK + b.g, b.sj,
                                                b.sk);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int DownloadAos(aither_gpu *h, const HostBlock &hb, double *dst, int SI, int SJ, int SK, int nc,
                const double *srcField, int originI, int originJ, int originK) {
  const size_t n = static_cast<size_t>(SI) * SJ * SK * nc;
  if (EnsureStage(h, n * sizeof(double))) return 1;
  const BlockDev &b = hb.dev;
  const long long cells = static_cast<long long>(SI) * SJ * SK;
  const int grid = static_cast<int>(std::min<long long>((cells + 255) / 256, 148 * 16));
  {
    ScopedLaunch sl(h, kFamLayout);
    SoaToAosKernel<<<grid, 256, 0, h->stream>>>(h->dStage, SI, SJ, SK, nc, srcField, b.fs,
                                                originI + b.lp, originJ + b.g, originK + b.g, b.sj,
                                                b.sk);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, h->dStage, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// Chunk length along k for a plane-marching kernel: every (column, chunk) pair is one thread
// block, `slots` blocks are resident at a time and a block's time is ~ its planes plus `overhead`
// prologue planes. The time of the launch is (number of waves) x (chunk + overhead): with 256
// planes, 128 columns and 148 slots a 52-plane chunk gives 4.3 -> 5 waves (270 units) where a
// 32-plane chunk gives 6.9 -> 7 waves (238 units): measured 1.22 -> 1.10 ms per DPLUR sweep.
int PickChunk(int nk, int cols, int slots, int overhead) {
  int best = std::max(1, std::min(nk, 8));
  double bestCost = 1e300;
  for (int c = std::min(nk, 8); c <= std::min(nk, 64); ++c) {
    const long long blocks = static_cast<long long>(cols) * ((nk + c - 1) / c);
    const long long waves = (blocks + slots - 1) / slots;
    // partially filled last wave still costs a full block; mild preference for longer chunks
    const double cost = static_cast<double>(waves) * (c + overhead) * (1.0 + 0.02 / c);
    if (cost < bestCost) {
      bestCost = cost;
      best = c;
    }
  }
  return best;
}

int SurfaceType(const aither_surface &s) {
  // ref: src/boundaryConditions.cpp:2424-2452
  if (s.imin == s.imax) return s.imax == 0 ? 1 : 2;
  if (s.jmin == s.jmax) return s.jmax == 0 ? 3 : 4;
  return s.kmax == 0 ? 5 : 6;
}

bool Supported(const aither_cfg &c, std::string *why) {
  if (c.numSpecies < 1 || c.numSpecies > 3) { *why = "numSpecies must be 1, 2 or 3 (kernels are instantiated for these counts)"; return false; }
  if (c.numSpecies > 1 && c.isBlockMatrix) { *why = "block-matrix solvers are built for one species (species rows of the thin-shear-layer Jacobian)"; return false; }
  if (c.numTurb != 0 || c.isRANS) {
    if (c.numTurb != 2 || !c.isRANS || !c.isViscous) { *why = "RANS needs numTurb = 2 and isViscous"; return false; }
    if (c.turbModel != AITHER_TURB_KW_WILCOX && c.turbModel != AITHER_TURB_SST) {
      *why = "turbulence model must be kOmegaWilcox2006 or sst2003"; return false;
    }
  }
  if (c.isViscous && c.viscRecon != 0 && c.viscRecon != 1) { *why = "unknown viscous face reconstruction"; return false; }
  if (c.isViscous && c.numGhosts < 2) { *why = "viscous fluxes need at least 2 ghost layers"; return false; }
  if (c.invFluxJac != AITHER_JAC_RUSANOV && c.invFluxJac != AITHER_JAC_APPROX_ROE) { *why = "unknown inviscid flux jacobian"; return false; }
  if (c.invFluxJac == AITHER_JAC_APPROX_ROE && (c.isViscous || c.isBlockMatrix)) {
    // RoeOffDiagonal receives (f1, dist) in swapped positions (reference src/fluxJacobian.cpp:226-228
    // vs :243-244) and divides the viscous spectral radius by f1 = 0 for laminar flow
    *why = "approximateRoe is built for inviscid runs with a scalar diagonal only (the reference's "
           "viscous branch divides by zero)";
    return false;
  }
  if (c.numGhosts < 1 || c.numGhosts > 3) { *why = "numGhosts must be 1..3"; return false; }
  {
    // the reconstruction stencil must fit the ghost shell (ref: src/input.cpp:1127-1144)
    const int need = c.recon == AITHER_RECON_CONSTANT ? 1 : (c.recon == AITHER_RECON_MUSCL ? 2 : 3);
    if (c.numGhosts < need) { *why = "numGhosts is smaller than the stencil of the face reconstruction (constant 1, MUSCL 2, WENO 3)"; return false; }
  }
  if (c.numBCStates < 0 || c.numBCStates > AITHER_MAX_BC_STATES) { *why = "numBCStates must be 0.." + std::to_string(AITHER_MAX_BC_STATES); return false; }
  for (int q = 0; q < c.numBCStates; ++q) {
    if (c.bcStates[q].isWallLaw && c.isViscous && !c.isRANS && c.numSpecies > 1) {
      *why = "the wall law in a laminar run is built for one species";
      return false;
    }
  }
  return true;
}

template <int NS, int NT, int RC, int LM, int FX>
void LaunchResidualOne(aither_gpu *h, HostBlock &hb, int implicitScalar, int fusePrep, double cfl) {
#ifdef AITHER_B200_LEGACY_RESIDUAL  // first-generation gather kernel, A/B builds only
  if (h->legacyKernels) {
    ResidualKernel<NS, NT, RC, LM, FX><<<hb.resGrid, hb.resBlock, 0, h->stream>>>(
        hb.dev, h->params, implicitScalar);
    return;
  }
#endif
  using S = ResSmem<NS, NT, RC>;
  if constexpr (S::bytes > 227 * 1024) {
    // 7 equations x WENO's 7-plane ring does not fit the 227 KB of shared memory: this one
    // combination takes the gather kernel (thread per cell, face fluxes shared per tile)
    ResidualKernel<NS, NT, RC, LM, FX><<<hb.resGrid, hb.resBlock, 0, h->stream>>>(
        hb.dev, h->params, implicitScalar);
    return;
  } else {
  auto kern = ResidualMarchKernel<NS, NT, RC, LM, FX>;
  static bool configured[64] = {false};  // per template instantiation and device
  EnsureSmemOptIn(kern, S::bytes, h->device, configured);
  kern<<<hb.marchGrid, dim3(kMI, kMJ, 1), S::bytes, h->stream>>>(hb.dev, h->params, hb.kChunk,
                                                                 implicitScalar, fusePrep, cfl);
  }
}

template <int NS, int NT>
int LaunchResidual(aither_gpu *h, HostBlock &hb, int fusePrep, double cfl) {
  const aither_cfg &c = h->cfg;
  const int implicitScalar = h->jac == kJacBlock ? 0 : 1;
  if (h->legacyKernels || h->jac == kJacBlock) fusePrep = 0;
  ScopedLaunch sl(h, kFamResidual);
#define RES(RC, LM, FX) LaunchResidualOne<NS, NT, RC, LM, FX>(h, hb, implicitScalar, fusePrep, cfl)
#define RES_FLUX(RC, LM)                         \
  do {                                           \
    if (c.invFlux == AITHER_FLUX_ROE) RES(RC, LM, AITHER_FLUX_ROE); \
    else RES(RC, LM, AITHER_FLUX_AUSM);          \
  } while (0)
  if (c.recon == AITHER_RECON_CONSTANT) {
    RES_FLUX(AITHER_RECON_CONSTANT, AITHER_LIMITER_NONE);
  } else if (c.recon == AITHER_RECON_MUSCL) {
    if (c.limiter == AITHER_LIMITER_NONE) RES_FLUX(AITHER_RECON_MUSCL, AITHER_LIMITER_NONE);
    else if (c.limiter == AITHER_LIMITER_VAN_ALBADA) RES_FLUX(AITHER_RECON_MUSCL, AITHER_LIMITER_VAN_ALBADA);
    else RES_FLUX(AITHER_RECON_MUSCL, AITHER_LIMITER_MINMOD);
  } else {
    RES_FLUX(AITHER_RECON_WENO, AITHER_LIMITER_NONE);
  }
#undef RES
#undef RES_FLUX
  return 0;
}

template <int NS, int NT, int MODE>
void LaunchImplicitMarch(aither_gpu *h, HostBlock &hb, const double *xin, double *xout,
                         int storeField) {
  constexpr size_t bytes = sizeof(double) * 2 * Ingr<NS, NT>::n * kIPC;
  auto kern = ImplicitMarchKernel<NS, NT, MODE>;
  static bool configured[64] = {false};
  EnsureSmemOptIn(kern, bytes, h->device, configured);
  kern<<<hb.marchGrid, dim3(kMI, kMJ, 1), bytes, h->stream>>>(hb.dev, h->params, xin, xout,
                                                             hb.kChunk, h->dPartials, storeField);
}

// `ordered`: one launch over the block's tile list (tiles next to a connected face first, each
// signalling h->dHaloSignal when done) instead of the plain 3-D grid
template <int NS, int NT, int MODE>
void LaunchImplicitTma(aither_gpu *h, HostBlock &hb, const double *xin, double *xout,
                       int storeField, bool ordered = false, double *stateOut = nullptr) {
  using T = ImplTma<NS, NT>;
  auto kern = ImplicitTmaKernel<NS, NT, MODE>;
  static bool configured[64] = {false};
  EnsureSmemOptIn(kern, T::bytes, h->device, configured);
  const BlockDev &b = hb.dev;
  const double *base = static_cast<const double *>(hb.alloc);
  const int fX = static_cast<int>((xin - base) / b.fs);
  const int fAi = static_cast<int>((b.fA[0] - base) / b.fs);
  const int fAj = static_cast<int>((b.fA[1] - base) / b.fs);
  const int fState = static_cast<int>((b.state - base) / b.fs);
  if (!ordered) {
    kern<<<hb.tmaGrid, dim3(kQI, kQJ, 1), T::bytes, h->stream>>>(
        hb.tmaMaps, b, h->params, xin, xout, fX, fAi, fAj, hb.tmaChunk, h->dPartials, storeField,
        nullptr, 0, 0, 0, nullptr, fState, stateOut);
    return;
  }
  kern<<<hb.nTmaBlocks, dim3(kQI, kQJ, 1), T::bytes, h->stream>>>(
      hb.tmaMaps, b, h->params, xin, xout, fX, fAi, fAj, hb.tmaChunk, h->dPartials, storeField,
      hb.dTmaTiles, hb.tmaGrid.x, hb.tmaGrid.y, hb.nTmaBoundaryTiles, h->dHaloSignal, fState,
      nullptr);
}

template <int NS, int NT>
void LaunchRansCell(const BlockDev &b, const Params &p, dim3 grid, cudaStream_t stream, bool block,
                    const EdgeSurf *surfs, int nsurf) {
  if (block) {
    if constexpr (NS == 1) {
      RansCellKernel<NS, NT, true><<<grid, dim3(32, 4, 1), 0, stream>>>(b, p, 0, surfs, nsurf);
    }
  } else {
    RansCellKernel<NS, NT, false><<<grid, dim3(32, 4, 1), 0, stream>>>(b, p, 1, surfs, nsurf);
  }
}

// inviscid part of the block diagonal, same reconstruction as the residual
template <int NS, int NT>
void LaunchBlockDiagInv(aither_gpu *h, HostBlock &hb) {
  const aither_cfg &c = h->cfg;
  const BlockDev &b = hb.dev;
  const dim3 grid((b.ni + 31) / 32, (b.nj + 3) / 4, b.nk), blk(32, 4, 1);
#define BDI(RC, LM) BlockDiagInvKernel<NS, NT, RC, LM><<<grid, blk, 0, h->stream>>>(b, h->params)
  if (c.recon == AITHER_RECON_CONSTANT) BDI(AITHER_RECON_CONSTANT, AITHER_LIMITER_NONE);
  else if (c.recon == AITHER_RECON_MUSCL) {
    if (c.limiter == AITHER_LIMITER_NONE) BDI(AITHER_RECON_MUSCL, AITHER_LIMITER_NONE);
    else if (c.limiter == AITHER_LIMITER_VAN_ALBADA) BDI(AITHER_RECON_MUSCL, AITHER_LIMITER_VAN_ALBADA);
    else BDI(AITHER_RECON_MUSCL, AITHER_LIMITER_MINMOD);
  } else BDI(AITHER_RECON_WENO, AITHER_LIMITER_NONE);
#undef BDI
}

// pencil lattice of a block for the LU-SGS wavefront kernels: ticket order along anti-diagonals
// and the ticket / progress counters
int EnsureWaveLattice(aither_gpu *h, HostBlock &hb, int TJ, int TK) {
  if (hb.waveTJ == TJ && hb.waveTK == TK) return 0;
  const BlockDev &b = hb.dev;
  const int nbJ = (b.nj + TJ - 1) / TJ, nbK = (b.nk + TK - 1) / TK;
  std::vector<int2> order;
  order.reserve(static_cast<size_t>(nbJ) * nbK);
  for (int s = 0; s <= nbJ + nbK - 2; ++s)
    for (int bk = std::max(0, s - (nbJ - 1)); bk <= std::min(s, nbK - 1); ++bk)
      order.push_back(make_int2(s - bk, bk));
  if (hb.dWaveOrder) cudaFree(hb.dWaveOrder);
  if (hb.dWaveSync) cudaFree(hb.dWaveSync);
  hb.dWaveOrder = nullptr;
  hb.dWaveSync = nullptr;
  hb.waveSyncBytes = sizeof(WaveSync) + sizeof(int) * order.size();
  CK(cudaMalloc(&hb.dWaveOrder, sizeof(int2) * order.size()));
  CK(cudaMalloc(&hb.dWaveSync, hb.waveSyncBytes));
  CK(cudaMemcpyAsync(hb.dWaveOrder, order.data(), sizeof(int2) * order.size(),
                     cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));  // `order` leaves scope
  hb.waveTJ = TJ;
  hb.waveTK = TK;
  hb.waveNbJ = nbJ;
  hb.wavePencils = static_cast<int>(order.size());
  return 0;
}

// LU-SGS half sweep as one persistent wavefront launch, eight lanes per cell (lusgs_wave.cuh):
// block matrices and approximateRoe off-diagonals
template <int NS, int NT, int JAC>
int LaunchLusgsWave(aither_gpu *h, HostBlock &hb, bool forward, int fullGS) {
  // 7 x 4 pencils: 224 + 32 = 256 threads, 255 registers
  constexpr int TJ = 7, TK = 4;
  const BlockDev &b = hb.dev;
  if (EnsureWaveLattice(h, hb, TJ, TK)) return 1;
  CK(cudaMemsetAsync(hb.dWaveSync, 0, hb.waveSyncBytes, h->stream));
  const int grid = std::min(hb.wavePencils, 148);
  if (forward)
    LusgsWaveKernel<NS, NT, true, JAC, TJ, TK><<<grid, TJ * TK * 8 + 32, 0, h->stream>>>(
        b, h->params, fullGS, hb.dWaveOrder, hb.wavePencils, hb.waveNbJ, hb.dWaveSync);
  else
    LusgsWaveKernel<NS, NT, false, JAC, TJ, TK><<<grid, TJ * TK * 8 + 32, 0, h->stream>>>(
        b, h->params, fullGS, hb.dWaveOrder, hb.wavePencils, hb.waveNbJ, hb.dWaveSync);
  return 0;
}

// scalar-diagonal LU-SGS (lusgs_pencil.cuh): per-iteration pack of the update-independent record
// into the plane-major workspace, then per half sweep the parallel ahead-sum pass and one
// persistent wavefront launch
PencilLattice LatticeOf(const HostBlock &hb) {
  PencilLattice L;
  L.nbJ = (hb.dev.nj + kPTJ - 1) / kPTJ;
  L.nbK = (hb.dev.nk + kPTK - 1) / kPTK;
  L.planesPer = hb.dev.ni + kPTJ + kPTK - 2;
  return L;
}
template <int NS, int NT, bool VISC>
int PackLusgsPencilV(aither_gpu *h, HostBlock &hb) {
  using R = PencilRec<NS, NT, VISC>;
  const BlockDev &b = hb.dev;
  const PencilLattice L = LatticeOf(hb);
  const long long slots = static_cast<long long>(L.nbJ) * L.nbK * L.planesPer * kPCells;
  const dim3 grid((b.ni + 31) / 32, (b.nj + 3) / 4, b.nk), blk(32, 4, 1);
  if (!hb.dWaveDyn) {
    CK(cudaMalloc(&hb.dWaveGeoLo, sizeof(double) * R::GN * slots));
    CK(cudaMalloc(&hb.dWaveGeoHi, sizeof(double) * R::GN * slots));
    CK(cudaMalloc(&hb.dWaveDyn, sizeof(double) * R::DN * slots));
    CK(cudaMalloc(&hb.dWaveAhead, sizeof(double) * R::AN * slots));
    // slots of clipped pencils / fill planes are copied with their plane: keep them finite
    CK(cudaMemsetAsync(hb.dWaveGeoLo, 0, sizeof(double) * R::GN * slots, h->stream));
    CK(cudaMemsetAsync(hb.dWaveGeoHi, 0, sizeof(double) * R::GN * slots, h->stream));
    CK(cudaMemsetAsync(hb.dWaveDyn, 0, sizeof(double) * R::DN * slots, h->stream));
    CK(cudaMemsetAsync(hb.dWaveAhead, 0, sizeof(double) * R::AN * slots, h->stream));
    ScopedLaunch sl(h, kFamLayout);
    WaveGeoKernel<R::GN><<<grid, blk, 0, h->stream>>>(b, L, h->cfg.isViscous, hb.dWaveGeoLo,
                                               hb.dWaveGeoHi);
  }
  // half sweeps hand the next one's ahead-sums on (AITHER_B200_LUSGS_CARRY=0: always the parallel
  // pass, for A/B runs); next to connected faces they are formed again after the exchange
  static const bool carryOff = getenv("AITHER_B200_LUSGS_CARRY") && atoi(getenv("AITHER_B200_LUSGS_CARRY")) == 0;
  hb.waveCarries = !carryOff;
  hb.waveFixup = false;
  for (int s = 0; s < 6; ++s) hb.waveFixup = hb.waveFixup || b.connFace[s] != nullptr;
  ScopedLaunch sl(h, kFamLusgsPack);
  WaveDynKernel<NS, NT, VISC><<<grid, blk, 0, h->stream>>>(b, h->params, L, hb.dWaveDyn);
  return 0;
}
// Euler runs (no turbulence equations, inviscid) take the records without the viscous slots
template <int NS, int NT>
int PackLusgsPencil(aither_gpu *h, HostBlock &hb) {
  if constexpr (NT == 0) {
    if (!h->cfg.isViscous) return PackLusgsPencilV<NS, NT, false>(h, hb);
  }
  return PackLusgsPencilV<NS, NT, true>(h, hb);
}
template <int NS, int NT, bool VISC>
int LaunchLusgsPencilV(aither_gpu *h, HostBlock &hb, bool forward, int fullGS) {
  using C = PencilCfg<NS, NT, VISC>;
  const BlockDev &b = hb.dev;
  const PencilLattice L = LatticeOf(hb);
  if (EnsureWaveLattice(h, hb, kPTJ, kPTK)) return 1;
  CK(cudaMemsetAsync(hb.dWaveSync, 0, hb.waveSyncBytes, h->stream));
  const int grid = std::min(hb.wavePencils, 148 * kPencilCtasPerSm);
  auto fwd = LusgsPencilKernel<NS, NT, true, VISC>;
  auto bwd = LusgsPencilKernel<NS, NT, false, VISC>;
  static bool cfgF[64] = {false}, cfgB[64] = {false};  // per device and template instantiation
  if (EnsureSmemOptIn(fwd, C::smemBytes, h->device, cfgF) ||
      EnsureSmemOptIn(bwd, C::smemBytes, h->device, cfgB))
    return Fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the LU-SGS wavefront");
  if (!hb.dWaveMailJ) {
    const size_t lines = static_cast<size_t>(hb.wavePencils) * L.planesPer;
    const size_t bytesJ = lines * kPTK * C::NI * sizeof(uint4), bytesK = lines * kPTJ * C::NI * sizeof(uint4);
    CK(cudaMalloc(&hb.dWaveMailJ, bytesJ));
    CK(cudaMalloc(&hb.dWaveMailK, bytesK));
    CK(cudaMemsetAsync(hb.dWaveMailJ, 0, bytesJ, h->stream));
    CK(cudaMemsetAsync(hb.dWaveMailK, 0, bytesK, h->stream));
  }
  if (++hb.waveTag == 0) ++hb.waveTag;  // 0 is what an untouched mailbox holds
  double *carry = hb.waveCarries ? hb.dWaveAhead : nullptr;
  // AITHER_B200_LUSGS_DBG=<file>: the time line of the pencils of the 6th forward launch
  static long long *dbg = nullptr;
  static int dbgCount = 0;
  const char *dbgFile = getenv("AITHER_B200_LUSGS_DBG");
  if (dbgFile && !dbg) {
    cudaMalloc(&dbg, sizeof(long long) * (8 * hb.wavePencils + 512));
    cudaMemset(dbg, 0, sizeof(long long) * (8 * hb.wavePencils + 512));
  }
  const bool record = dbg && forward && ++dbgCount == 6;
  static const int dbgFlags = getenv("AITHER_B200_LUSGS_DBGFLAGS") ? atoi(getenv("AITHER_B200_LUSGS_DBGFLAGS")) : 0;
  if (forward)
    fwd<<<grid, C::threads, C::smemBytes, h->stream>>>(
        b, h->params, L, fullGS, hb.dWaveDyn, hb.dWaveGeoLo, hb.dWaveAhead, hb.dWaveOrder,
        hb.wavePencils, hb.dWaveSync, hb.dWaveMailJ, hb.dWaveMailK, hb.waveTag, carry,
        record ? dbg : nullptr, dbgFlags);
  else
    bwd<<<grid, C::threads, C::smemBytes, h->stream>>>(
        b, h->params, L, fullGS, hb.dWaveDyn, hb.dWaveGeoHi, hb.dWaveAhead, hb.dWaveOrder,
        hb.wavePencils, hb.dWaveSync, hb.dWaveMailJ, hb.dWaveMailK, hb.waveTag, carry, nullptr, dbgFlags);
  if (record) {
    std::vector<long long> hbuf(8 * hb.wavePencils + 512);
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(hbuf.data(), dbg, sizeof(long long) * hbuf.size(), cudaMemcpyDeviceToHost);
    if (FILE *f = fopen(dbgFile, "w")) {
      fprintf(f, "# ticket bJ bK smid | ns since the first ticket: drawn, plane 0, 32, 64, 96, 128 reached, last plane done\n");
      for (int t = 0; t < hb.wavePencils; ++t) {
        fprintf(f, "%d %lld %lld %lld |", t, hbuf[8 * t + 7] % 1000, hbuf[8 * t + 7] / 1000 % 1000,
                hbuf[8 * t + 7] / 1000000);
        for (int k = 0; k < 7; ++k) fprintf(f, " %lld", hbuf[8 * t + k] ? hbuf[8 * t + k] - hbuf[0] : -1LL);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return 0;
}
template <int NS, int NT>
int LaunchLusgsPencil(aither_gpu *h, HostBlock &hb, bool forward, int fullGS) {
  if constexpr (NT == 0) {
    if (!h->cfg.isViscous) return LaunchLusgsPencilV<NS, NT, false>(h, hb, forward, fullGS);
  }
  return LaunchLusgsPencilV<NS, NT, true>(h, hb, forward, fullGS);
}
// the ahead-side sums (old update) of a half sweep in one parallel pass, before the wavefront
// `fixupOnly`: the previous half sweep left this one's ahead-sums behind; only the cells whose
// ahead-neighbour is a ghost cell across a connection (exchanged since) are formed again
template <int NS, int NT, bool VISC>
int LaunchLusgsAheadV(aither_gpu *h, HostBlock &hb, bool forward, bool fixupOnly) {
  const BlockDev &b = hb.dev;
  const PencilLattice L = LatticeOf(hb);
  const dim3 ablk(32, 4, 1);
  auto launch = [&](int i0, int j0, int k0, int i1, int j1, int k1) {
    const dim3 agrid((i1 - i0 + 31) / 32, (j1 - j0 + 3) / 4, k1 - k0);
    ScopedLaunch sa(h, kFamLusgsAhead);
    if (forward)
      LusgsAheadKernel<NS, NT, true, VISC><<<agrid, ablk, 0, h->stream>>>(b, h->params, L, hb.dWaveAhead,
                                                                         i0, j0, k0, i1, j1);
    else
      LusgsAheadKernel<NS, NT, false, VISC><<<agrid, ablk, 0, h->stream>>>(b, h->params, L, hb.dWaveAhead,
                                                                          i0, j0, k0, i1, j1);
  };
  if (!fixupOnly) {
    launch(0, 0, 0, b.ni, b.nj, b.nk);
    return 0;
  }
  // the ahead side of a forward sweep is the upper side (surfaces 2, 4, 6), of a backward one the lower
  if (b.connFace[forward ? 1 : 0]) { const int i = forward ? b.ni - 1 : 0; launch(i, 0, 0, i + 1, b.nj, b.nk); }
  if (b.connFace[forward ? 3 : 2]) { const int j = forward ? b.nj - 1 : 0; launch(0, j, 0, b.ni, j + 1, b.nk); }
  if (b.connFace[forward ? 5 : 4]) { const int k = forward ? b.nk - 1 : 0; launch(0, 0, k, b.ni, b.nj, k + 1); }
  return 0;
}
template <int NS, int NT>
int LaunchLusgsAhead(aither_gpu *h, HostBlock &hb, bool forward, bool fixupOnly) {
  if constexpr (NT == 0) {
    if (!h->cfg.isViscous) return LaunchLusgsAheadV<NS, NT, false>(h, hb, forward, fixupOnly);
  }
  return LaunchLusgsAheadV<NS, NT, true>(h, hb, forward, fixupOnly);
}

int ZeroResult(aither_gpu *h, int slot) {
  CK(cudaMemsetAsync(h->dResults + slot, 0, sizeof(IterResult), h->stream));
  return 0;
}

int EnsureResults(aither_gpu *h, int n) {
  if (n <= h->resultsCap) return 0;
  if (h->dResults) cudaFree(h->dResults);
  if (h->hResults) cudaFreeHost(h->hResults);
  h->dResults = nullptr;
  h->hResults = nullptr;
  CK(cudaMalloc(&h->dResults, sizeof(IterResult) * n));
  CK(cudaMallocHost(&h->hResults, sizeof(IterResult) * n));
  h->resultsCap = n;
  return 0;
}

// ---- phases (all asynchronous on h->stream) ---------------------------------------------------
// The exchange itself, issued on the communication stream; the caller orders it against the
// compute stream (ExchangeBegin / ExchangeEnd below).
int ExchangeOnComm(aither_gpu *h, int which, cudaStream_t st = nullptr) {
  if (st == nullptr) st = h->commStream;
  // ref: src/gridLevel.cpp:297-312 (state), src/utility.cpp:400-423 (implicit update),
  // src/procBlock.cpp:3064-3085 (eddy viscosity + f1 + f2: three contiguous fields; velocity gradient)
  // only the state of viscous runs needs the edge ghost cells (Green-Gauss stencils); everything
  // else is read face-normal and takes the single-level plan
  // (and the wall distance of the set-up, swapped like any slice: src/gridLevel.cpp:261-281)
  HaloPlan &plan = ((which == kHaloState && h->stateNeedsEdges) || which == kHaloWallDist)
                       ? h->halo
                       : ((which == kHaloUpdate && h->updateOneLayer) ? h->haloUpdate : h->haloFace);
  const int total = which == kHaloTurb ? 3 : (which == kHaloVelGrad ? 9 : (which == kHaloWallDist ? 1 : h->neq));
  for (int done = 0; done < total;) {
    const int nc = std::min(total - done, h->neq);  // the plan's buffers hold neq components
    HaloFields f;
    for (size_t bb = 0; bb < h->blocks.size(); ++bb) {
      const BlockDev &b = h->blocks[bb].dev;
      double *base = which == kHaloState ? b.state
                     : which == kHaloUpdate ? b.x
                     : which == kHaloTurb ? b.eddyVisc
                     : which == kHaloWallDist ? b.wallDist : b.velGrad;
      f.base[bb] = base + static_cast<long long>(done) * b.fs;
      f.fs[bb] = b.fs;
    }
    ScopedLaunch sl(h, kFamHalo, st);
    h->launches--;  // ScopedLaunch counts one; the exchange counts its own kernels below
    h->famLaunches[kFamHalo]--;
    if (HaloExchange(plan, f, nc, st, &h->launches, &h->famLaunches[kFamHalo]))
      return Fail(HaloError());
    done += nc;
  }
  return 0;
}
// what the exchange packs has been written by everything issued on the compute stream so far
int ExchangeBegin(aither_gpu *h) {
  CK(cudaEventRecord(h->evCompute, h->stream));
  CK(cudaStreamWaitEvent(h->commStream, h->evCompute, 0));
  return 0;
}
// ... and what is issued on the compute stream from here on sees the ghost cells it wrote
int ExchangeEnd(aither_gpu *h) {
  CK(cudaEventRecord(h->evExchanged, h->commStream));
  CK(cudaStreamWaitEvent(h->stream, h->evExchanged, 0));
  return 0;
}
// a ghost exchange in program order of the compute stream (on the compute stream itself unless
// the overlapped sweeps use the communication stream: NCCL wants one stream per communicator)
int Exchange(aither_gpu *h, int which) {
  if (h->halo.nConn == 0) return 0;
  if (!h->haloOverlap) return ExchangeOnComm(h, which, h->stream);
  if (ExchangeBegin(h) || ExchangeOnComm(h, which)) return 1;
  return ExchangeEnd(h);
}

template <int NS, int NT>
int PhaseBoundaryConditionsT(aither_gpu *h) {
  // ref: src/gridLevel.cpp:287-319
  for (auto &hb : h->blocks) {
    if (hb.bcThreads == 0) continue;
    ScopedLaunch sl(h, kFamBc);
    const int grid = static_cast<int>((hb.bcThreads + 127) / 128);
    if (hb.dPatchMach)
      PatchMachKernel<NS, NT><<<hb.nBcSurfs, 256, 0, h->stream>>>(hb.dev, h->params, hb.dSurfs,
                                                                  h->dBcStates, hb.dPatchMach);
    BcKernel<NS, NT><<<grid, 128, 0, h->stream>>>(hb.dev, h->params, hb.dSurfs, hb.nBcSurfs,
                                               h->dBcStates, hb.bcThreads, hb.dPatchMach);
  }
  CK(cudaGetLastError());
  if (Exchange(h, kHaloState)) return 1;
  if (h->cfg.isViscous || h->nonreflecting) {
    // edge ghost cells: read by the gradient stencils only (viscous fluxes, or the gradient pass
    // an Euler run makes for its non-reflecting BCs; ref src/gridLevel.cpp:314-318)
    for (auto &hb : h->blocks) {
      ScopedLaunch sl(h, kFamViscGhost);
      const int n = 4 * (hb.dev.ni + hb.dev.nj + hb.dev.nk);
      EdgeKernel<NS, NT, false><<<(n + 127) / 128, 128, 0, h->stream>>>(
          hb.dev, h->params, hb.dEdgeSurfs, hb.nEdgeSurfs, h->dBcStates);
    }
    CK(cudaGetLastError());
  }
  return 0;
}

template <int NS, int NT>
int PhaseResidualT(aither_gpu *h, int fusePrep, double cfl) {
  const bool block = h->jac == kJacBlock;
  for (auto &hb : h->blocks) {
    LaunchResidual<NS, NT>(h, hb, fusePrep, cfl);
    if constexpr (NS == 1) {
      if (block) {
        ScopedLaunch sl(h, kFamResidual);
        LaunchBlockDiagInv<NS, NT>(h, hb);
      }
    }
  }
  CK(cudaGetLastError());
  if (!h->cfg.isViscous) {
    // Euler run with a non-reflecting BC: the reference's gradient-only pass
    // (CalcGradsI/J/K, src/procBlock.cpp:6138-6146, :5790-5945) leaves the same cell averages
    if (h->nonreflecting) {
      for (auto &hb : h->blocks) {
        ScopedLaunch sl(h, kFamResidual);
        const BlockDev &b = hb.dev;
        const dim3 grid((b.ni + 31) / 32, (b.nj + 3) / 4, b.nk);
        CellGradKernel<NS, NT><<<grid, dim3(32, 4, 1), 0, h->stream>>>(b, 1);
      }
      CK(cudaGetLastError());
    }
    return 0;
  }
  // ref: src/procBlock.cpp:6125-6137
  for (auto &hb : h->blocks) {
    const BlockDev &b = hb.dev;
    if (hb.bcThreads > 0) {
      ScopedLaunch sl(h, kFamViscGhost);
      const int grid = static_cast<int>((hb.bcThreads + 127) / 128);
      ViscousWallKernel<NS, NT><<<grid, 128, 0, h->stream>>>(b, h->params, hb.dSurfs, hb.nBcSurfs,
                                                          h->dBcStates, hb.bcThreads);
    }
    {
      ScopedLaunch sl(h, kFamViscGhost);
      const int n = 4 * (b.ni + b.nj + b.nk);
      EdgeKernel<NS, NT, true><<<(n + 127) / 128, 128, 0, h->stream>>>(
          b, h->params, hb.dEdgeSurfs, hb.nEdgeSurfs, h->dBcStates);
    }
    {
      ScopedLaunch sl(h, kFamViscGhost);
      const dim3 grid((b.ni + 2 * b.g + 31) / 32, (b.nj + 2 * b.g + 7) / 8, b.nk + 2 * b.g);
      AuxKernel<NS, NT><<<grid, dim3(32, 8, 1), 0, h->stream>>>(b, h->params);
    }
    if (NT > 0 || block || NS > 1 || b.wallVars != nullptr) {
      // RANS, multi-species, block matrix and / or wall-law walls: viscous (+ turbulent) fluxes, cell averages, spectral radii,
      // source terms and the thin-shear-layer Jacobians, per cell
      ScopedLaunch sl(h, kFamViscFlux);
      const dim3 grid((b.ni + 31) / 32, (b.nj + 3) / 4, b.nk);
      LaunchRansCell<NS, NT>(b, h->params, grid, h->stream, block, hb.dEdgeSurfs, hb.nEdgeSurfs);
      if (b.pressGrad)  // non-reflecting BCs of the next iteration read the pressure gradient
        CellGradKernel<NS, NT><<<grid, dim3(32, 4, 1), 0, h->stream>>>(b, 0);
      continue;
    }
    {
      ScopedLaunch sl(h, kFamViscFlux);
      const dim3 grid((b.ni + 1 + 31) / 32, (b.nj + 1 + 7) / 8, b.nk + 1);
      ViscFaceKernel<NS, NT><<<grid, dim3(32, 8, 1), 0, h->stream>>>(b, h->params, b.xalt, b.x);
    }
    {
      ScopedLaunch sl(h, kFamViscFlux);
      ViscAccumKernel<NS, NT><<<hb.cellGrid, hb.cellBlock, 0, h->stream>>>(b, h->params, b.xalt, b.x,
                                                                         1);
    }
    if (b.pressGrad) {  // ... and the velocity gradient, which this path does not keep
      ScopedLaunch sl(h, kFamViscFlux);
      const dim3 grid((b.ni + 31) / 32, (b.nj + 3) / 4, b.nk);
      CellGradKernel<NS, NT><<<grid, dim3(32, 4, 1), 0, h->stream>>>(b, 1);
    }
  }
  CK(cudaGetLastError());
  // eddy viscosity and blending functions of the cells across connections
  // (gridLevel::SwapEddyViscAndGradients / SwapTurbVars, ref src/gridLevel.cpp:386-392)
  if (NT > 0 && Exchange(h, kHaloTurb)) return 1;
  // the block off-diagonals read the neighbour's velocity gradient (SwapEddyViscAndGradientSlice)
  if (block && Exchange(h, kHaloVelGrad)) return 1;
  return 0;
}

template <int NS, int NT>
int PhasePrepT(aither_gpu *h, double cfl, int bits) {
  for (auto &hb : h->blocks) {
    ScopedLaunch sl(h, kFamPrep);
    if (NS == 1 && h->jac == kJacBlock) {
      if constexpr (NS == 1) {
        const BlockDev &b = hb.dev;
        const dim3 grid((b.ni + 31) / 32, (b.nj + 3) / 4, b.nk);
        PrepBlockKernel<NS, NT><<<grid, dim3(32, 4, 1), 0, h->stream>>>(b, h->params, cfl, bits,
                                                                       h->dFlag);
      }
    } else {
      PrepKernel<NS, NT><<<hb.cellGrid, hb.cellBlock, 0, h->stream>>>(hb.dev, h->params, cfl, bits);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

int SwapUpdate(aither_gpu *h) {
  // ref: src/linearSolver.cpp:190-193, src/utility.cpp:400-423
  return Exchange(h, kHaloUpdate);
}

template <int NS, int NT, int JAC>
int PhaseRelaxJ(aither_gpu *h, int sweeps, int slot) {
  constexpr bool kCell = NT > 0 || JAC != kJacScalar || NS > 1;  // cell-parallel implicit kernels
  const bool fullGSAlways = h->cfg.matrixRequiresInit != 0;
  if constexpr (JAC == kJacScalar) {
    if (h->cfg.solver != AITHER_SOLVER_DPLUR && h->lusgsWave && sweeps > 0)
      for (auto &hb : h->blocks)
        if (PackLusgsPencil<NS, NT>(h, hb)) return 1;
  }
  // DPLUR with the TMA sweep and connections: every sweep runs the tiles next to a connected block
  // face first; their part of the new update is then exchanged on the communication stream while
  // the interior tiles are computed. The exchange after the last sweep is the one the matrix
  // residual needs. (ref: the reference swaps, blocking, before every sweep: src/linearSolver.cpp:
  // 190-193, src/utility.cpp:400-423 -- same data, same order, the wait is what moves.)
  bool overlapped = false;
  if constexpr (!kCell) {
    overlapped = h->cfg.solver == AITHER_SOLVER_DPLUR && h->haloOverlap && h->tmaImplicit &&
                 !h->legacyKernels && h->halo.nConn > 0 && sweeps > 0;
    for (auto &hb : h->blocks) overlapped = overlapped && hb.dTmaTiles != nullptr;
    if (overlapped) {
      if (SwapUpdate(h)) return 1;  // ghost cells of x0
      for (int s = 0; s < sweeps; ++s) {
        for (auto &hb : h->blocks) {
          ScopedLaunch sl(h, kFamDplur);
          LaunchImplicitTma<NS, NT, kModeDplur>(h, hb, hb.dev.x, hb.dev.xalt, 0, true);
          h->haloTarget += static_cast<unsigned int>(hb.nTmaBoundaryTiles);
          std::swap(hb.dev.x, hb.dev.xalt);
        }
        // the exchange of the new update starts when every boundary tile of this sweep has
        // signalled (the counter only grows: no reset, no ambiguity between sweeps) ...
        if (StreamWaitValueFn()(h->commStream, reinterpret_cast<CUdeviceptr>(h->dHaloSignal),
                                h->haloTarget, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
          return Fail("cuStreamWaitValue32 failed");
        if (ExchangeOnComm(h, kHaloUpdate)) return 1;
        if (ExchangeEnd(h)) return 1;  // ... and the next kernel on the compute stream waits for it
      }
      CK(cudaGetLastError());
    }
  }
  for (int s = 0; s < sweeps && !overlapped; ++s) {
    if (SwapUpdate(h)) return 1;
    if (h->cfg.solver == AITHER_SOLVER_DPLUR) {
      for (auto &hb : h->blocks) {
        {
          ScopedLaunch sl(h, kFamDplur);
          if (h->legacyKernels || kCell) {
            if (JAC == kJacScalar)
              DplurKernel<NS, NT, JAC><<<hb.cellGrid, hb.cellBlock, 0, h->stream>>>(
                  hb.dev, h->params, hb.dev.x, hb.dev.xalt);
            else
              DplurKernel<NS, NT, JAC><<<hb.cell128Grid, dim3(32, 4, 1), 0, h->stream>>>(
                  hb.dev, h->params, hb.dev.x, hb.dev.xalt);
          } else if constexpr (!kCell) {
            if (h->tmaImplicit) LaunchImplicitTma<NS, NT, kModeDplur>(h, hb, hb.dev.x, hb.dev.xalt, 0);
            else LaunchImplicitMarch<NS, NT, kModeDplur>(h, hb, hb.dev.x, hb.dev.xalt, 0);
          }
        }
        std::swap(hb.dev.x, hb.dev.xalt);
      }
    } else {
      const int fullGS = (s > 0 || fullGSAlways) ? 1 : 0;
      // one launch per i + j + k hyperplane; the ~(ni + nj + nk) launches of a half sweep are
      // captured once per block into a CUDA graph and replayed (launch-bound otherwise)
      auto halfSweep = [&](HostBlock &hb, bool forward) -> int {
        const BlockDev &b = hb.dev;
        const dim3 blk(16, 8);
        const dim3 grid((b.nj + 15) / 16, (b.nk + 7) / 8);
        const int last = b.ni + b.nj + b.nk - 3;
        // eight lanes per cell (LusgsPlaneSplitKernel): 32 cells per 256-thread block
        const int splitGrid = (b.nj * b.nk + 31) / 32;
        auto launchAll = [&]() {
          for (int n = 0; n <= last; ++n) {
            const int pl = forward ? n : last - n;
            if (h->lusgsSplit) {
              if (forward)
                LusgsPlaneSplitKernel<NS, NT, true, JAC><<<splitGrid, 256, 0, h->stream>>>(
                    b, h->params, pl, fullGS);
              else
                LusgsPlaneSplitKernel<NS, NT, false, JAC><<<splitGrid, 256, 0, h->stream>>>(
                    b, h->params, pl, fullGS);
            } else if (forward) {
              LusgsPlaneKernel<NS, NT, true, JAC><<<grid, blk, 0, h->stream>>>(b, h->params, pl,
                                                                              fullGS);
            } else {
              LusgsPlaneKernel<NS, NT, false, JAC><<<grid, blk, 0, h->stream>>>(b, h->params, pl,
                                                                               fullGS);
            }
          }
        };
        if constexpr (JAC == kJacScalar) {
          // the ahead-sums of this half sweep: left behind by the previous half sweep of this
          // iteration, unless ghost cells changed in between (connections) or there was none
          // (blocks with connections: the cells next to a connected face on the ahead side are
          // formed again, their ghost neighbours have been exchanged since)
          const bool carried = hb.waveCarries && !(s == 0 && forward);
          if (h->lusgsWave && fullGS && (!carried || hb.waveFixup) &&
              LaunchLusgsAhead<NS, NT>(h, hb, forward, carried))
            return 1;
        }
        ScopedLaunch sl(h, kFamLusgs);  // one timing record per half sweep
        if (h->lusgsWave) {
          if constexpr (JAC == kJacScalar) return LaunchLusgsPencil<NS, NT>(h, hb, forward, fullGS);
          else return LaunchLusgsWave<NS, NT, JAC>(h, hb, forward, fullGS);
        }
        h->launches += last;
        h->famLaunches[kFamLusgs] += last;
        if (!h->lusgsGraphs) {
          launchAll();
          return 0;
        }
        cudaGraphExec_t &ex = hb.lusgsGraph[forward ? 0 : 1][fullGS];
        if (!ex) {
          cudaGraph_t graph = nullptr;
          CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
          launchAll();
          CK(cudaStreamEndCapture(h->stream, &graph));
          CK(cudaGraphInstantiate(&ex, graph, 0));
          CK(cudaGraphDestroy(graph));
        }
        CK(cudaGraphLaunch(ex, h->stream));
        return 0;
      };
      for (auto &hb : h->blocks)
        if (halfSweep(hb, true)) return 1;
      if (SwapUpdate(h)) return 1;
      for (auto &hb : h->blocks)
        if (halfSweep(hb, false)) return 1;
    }
  }
  CK(cudaGetLastError());
  if (!overlapped && SwapUpdate(h)) return 1;
  // matrix residual and its norm (ref: src/linearSolver.cpp:92-109, src/mgSolution.cpp:198-206)
  for (auto &hb : h->blocks) {
    int nPartials = hb.nCellBlocks;
    {
      ScopedLaunch sl(h, kFamAxmb);
      if (h->legacyKernels || kCell) {
        if (JAC == kJacScalar) {
          AxmbKernel<NS, NT, JAC><<<hb.cellGrid, hb.cellBlock, 0, h->stream>>>(
              hb.dev, h->params, h->dPartials, h->keepMatrixResid ? 1 : 0);
        } else {
          AxmbKernel<NS, NT, JAC><<<hb.cell128Grid, dim3(32, 4, 1), 0, h->stream>>>(
              hb.dev, h->params, h->dPartials, h->keepMatrixResid ? 1 : 0);
          nPartials = hb.cell128Grid.x * hb.cell128Grid.y * hb.cell128Grid.z;
        }
      } else if constexpr (!kCell) {
        if (h->tmaImplicit) {
          double *stateOut = h->fuseUpdate ? hb.dev.stateAlt : nullptr;
          LaunchImplicitTma<NS, NT, kModeAxmb>(h, hb, hb.dev.x, nullptr, h->keepMatrixResid ? 1 : 0,
                                               false, stateOut);
          if (stateOut) h->stateFused = true;
          nPartials = hb.nTmaBlocks;
        } else {
          LaunchImplicitMarch<NS, NT, kModeAxmb>(h, hb, hb.dev.x, nullptr,
                                                 h->keepMatrixResid ? 1 : 0);
          nPartials = hb.nMarchBlocks;
        }
      }
    }
    {
      ScopedLaunch sl(h, kFamReduce);
      FinalizeSumKernel<<<1, kFinalThreads, 0, h->stream>>>(h->dPartials, nPartials, 1,
                                                            &h->dResults[slot].matrixSumSq);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

template <int NS, int NT>
int PhaseRelaxT(aither_gpu *h, int sweeps, int slot) {
  if constexpr (NS == 1) {
    if (h->jac == kJacBlock) return PhaseRelaxJ<NS, NT, kJacBlock>(h, sweeps, slot);
  }
  if constexpr (NT == 0) {
    if (h->jac == kJacRoe) return PhaseRelaxJ<NS, NT, kJacRoe>(h, sweeps, slot);
  }
  return PhaseRelaxJ<NS, NT, kJacScalar>(h, sweeps, slot);
}

template <int NS, int NT>
int PhaseUpdateT(aither_gpu *h, int slot, int mm) {
  const int nl = std::max(1, h->cfg.nonlinearIterations);
  for (auto &hb : h->blocks) {
    {
      ScopedLaunch sl(h, kFamUpdate);
      if (h->stateFused) {
        // the matrix-residual pass has advanced the state into the alternate buffer: residual
        // norms only, then the two buffers change roles
        UpdateKernel<NS, NT, true, false><<<hb.updGrid, hb.cellBlock, 0, h->stream>>>(
            hb.dev, h->params, h->dPartials, h->dLinfPartials);
        std::swap(hb.dev.state, hb.dev.stateAlt);
        hb.ghostsInAlt = true;
      } else {
        UpdateKernel<NS, NT><<<hb.updGrid, hb.cellBlock, 0, h->stream>>>(hb.dev, h->params,
                                                                        h->dPartials,
                                                                        h->dLinfPartials);
      }
    }
    {
      ScopedLaunch sl(h, kFamReduce);
      FinalizeSumKernel<<<h->neq, kFinalThreads, 0, h->stream>>>(h->dPartials, hb.nUpdBlocks,
                                                                 h->neq, h->dResults[slot].l2);
    }
    {
      ScopedLaunch sl(h, kFamReduce);
      FinalizeLinfKernel<<<1, kFinalThreads, 0, h->stream>>>(h->dLinfPartials, hb.nUpdBlocks,
                                                             hb.dev, h->neq, &h->dResults[slot]);
    }
    // U^(n-1) <- U^n after the last nonlinear iteration of a multilevel scheme
    // (ref: src/gridLevel.cpp:427-430)
    if (h->cfg.isMultilevelTime && mm == nl - 1) {
      CK(cudaMemcpyAsync(hb.dev.consNm1, hb.dev.consN, sizeof(double) * h->neq * hb.dev.fs,
                         cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// Equation-set dispatch: one or three species, laminar / Euler (NT = 0) or two-equation RANS
// (NT = 2). The phase templates of one equation set are instantiated in their own translation
// unit -- this same file compiled with -DAITHER_EQ_TU=<10|12|20|22|30|32> -- and reached through a
// table, so the library builds in parallel (__graft_entry__.build); with neither AITHER_EQ_TU nor
// AITHER_MAIN_TU defined the file is the whole library (one slow translation unit).
const AitherEqOps *EqOpsFor(const aither_gpu *h);
#define EQ_DISPATCH(h, FN, ...) (EqOpsFor(h)->FN(__VA_ARGS__))
int PhaseBoundaryConditions(aither_gpu *h) {
  for (auto &hb : h->blocks) hb.ghostsInAlt = false;  // the fill writes the current buffer's ghost cells
  return EQ_DISPATCH(h, PhaseBoundaryConditionsT, h);
}
int PhaseResidual(aither_gpu *h, int fusePrep = 0, double cfl = 0.0) {
  return EQ_DISPATCH(h, PhaseResidualT, h, fusePrep, cfl);
}
int PhasePrep(aither_gpu *h, double cfl, int bits) { return EQ_DISPATCH(h, PhasePrepT, h, cfl, bits); }
int PhaseRelax(aither_gpu *h, int sweeps, int slot) { return EQ_DISPATCH(h, PhaseRelaxT, h, sweeps, slot); }
int PhaseUpdate(aither_gpu *h, int slot, int mm) {
  h->stateMovedSinceStore = true;
  return EQ_DISPATCH(h, PhaseUpdateT, h, slot, mm);
}

template <int NS, int NT>
int StoreOldT(aither_gpu *h, int copyNm1) {
  for (auto &hb : h->blocks) {
    ScopedLaunch sl(h, kFamStore);
    StoreOldKernel<NS, NT><<<hb.cellGrid, hb.cellBlock, 0, h->stream>>>(hb.dev, h->params, copyNm1);
  }
  return 0;
}
// temperature and viscosity from the current state (gridLevel::AuxillaryAndWidths before the first
// iteration, ref src/gridLevel.cpp:433-438: the viscous-wall omega BC reads the stored viscosity)
template <int NS, int NT>
int InitAuxT(aither_gpu *h, int blk) {
  HostBlock &hb = h->blocks[blk];
  const BlockDev &b = hb.dev;
  ScopedLaunch sl(h, kFamViscGhost);
  const dim3 grid((b.ni + 2 * b.g + 31) / 32, (b.nj + 2 * b.g + 7) / 8, b.nk + 2 * b.g);
  AuxKernel<NS, NT><<<grid, dim3(32, 8, 1), 0, h->stream>>>(b, h->params);
  return 0;
}

// one function-file variable of block `blk` into the staging buffer (device), physical cells
template <int NS, int NT>
int OutputVarT(aither_gpu *h, int blk, int var, int species, double scale, double *dDst) {
  HostBlock &hb = h->blocks[blk];
  ScopedLaunch sl(h, kFamLayout);
  OutputVarKernel<NS, NT><<<hb.cellGrid, hb.cellBlock, 0, h->stream>>>(hb.dev, h->params, var,
                                                                      species, scale, dDst);
  return 0;
}

long long TotalPaddedSize(const aither_gpu *h) {
  long long t = 0;
  for (auto &hb : h->blocks) t += hb.paddedCells * h->neq;
  return t;
}

int IterateAsync(aither_gpu *h, double cfl, int slot, int mm) {
  if (ZeroResult(h, slot)) return 1;
  if (PhaseBoundaryConditions(h)) return 1;
  // inviscid: time step, diagonal, right-hand side and x0 ride in the residual kernel's epilogue
  const bool fuse = !h->legacyKernels && h->fusePrep && !h->cfg.isViscous && h->jac != kJacBlock;
  if (PhaseResidual(h, fuse ? 1 : 0, cfl)) return 1;
  if (!fuse && PhasePrep(h, cfl, kPrepDt | kPrepDiag | kPrepInit)) return 1;
  // the matrix-residual pass of the TMA path advances the state as well (A/B:
  // AITHER_B200_FUSE_UPDATE=0 leaves it to the update kernel)
  h->fuseUpdate = h->tmaImplicit && !h->legacyKernels && getenv("AITHER_B200_FUSE_UPDATE") == nullptr;
  h->stateFused = false;
  const int rcRelax = PhaseRelax(h, h->cfg.matrixSweeps, slot);
  h->fuseUpdate = false;
  if (rcRelax) return 1;
  const int rcUpd = PhaseUpdate(h, slot, mm);
  h->stateFused = false;
  return rcUpd;
}

int CheckFlag(aither_gpu *h) {
  if (h->jac != kJacBlock) return 0;
  int flag = 0;
  CK(cudaMemcpyAsync(&flag, h->dFlag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (flag != 0)  // the reference exits (src/matrix.cpp:83-86); here the call fails
    return Fail("singular diagonal block in the Gauss-Jordan inverse");
  return 0;
}

int FetchResults(aither_gpu *h, int n) {
  CK(cudaMemcpyAsync(h->hResults, h->dResults, sizeof(IterResult) * n, cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CheckFlag(h);
}

void FreeAll(aither_gpu *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (unsigned char *p : h->p2pPeers)
    if (p) cudaIpcCloseMemHandle(p);
  h->p2pPeers.clear();
  if (h->p2pArena) cudaFree(h->p2pArena);
  h->p2pArena = nullptr;
  for (auto &hb : h->blocks) {
    if (hb.alloc) cudaFree(hb.alloc);
    for (auto &row : hb.lusgsGraph)
      for (auto &ex : row)
        if (ex) cudaGraphExecDestroy(ex);
    if (hb.dTmaTiles) cudaFree(hb.dTmaTiles);
    if (hb.dWaveOrder) cudaFree(hb.dWaveOrder);
    if (hb.dWaveSync) cudaFree(hb.dWaveSync);
    if (hb.dWaveGeoLo) cudaFree(hb.dWaveGeoLo);
    if (hb.dWaveGeoHi) cudaFree(hb.dWaveGeoHi);
    if (hb.dWaveDyn) cudaFree(hb.dWaveDyn);
    if (hb.dWaveAhead) cudaFree(hb.dWaveAhead);
    if (hb.dWaveMailJ) cudaFree(hb.dWaveMailJ);
    if (hb.dWaveMailK) cudaFree(hb.dWaveMailK);
    if (hb.dSurfs) cudaFree(hb.dSurfs);
    if (hb.dEdgeSurfs) cudaFree(hb.dEdgeSurfs);
    if (hb.dWallVars) cudaFree(hb.dWallVars);
    if (hb.dPatchMach) cudaFree(hb.dPatchMach);
    if (hb.dToCoarse) cudaFree(hb.dToCoarse);
    if (hb.dChildren) cudaFree(hb.dChildren);
    if (hb.dVolFac) cudaFree(hb.dVolFac);
    if (hb.dProlong) cudaFree(hb.dProlong);
    if (hb.dSavedX) cudaFree(hb.dSavedX);
    if (hb.dNodes) cudaFree(hb.dNodes);
    if (hb.dSavedDiag) cudaFree(hb.dSavedDiag);
    for (auto &p : hb.dConnFace)
      if (p) cudaFree(p);
  }
  HaloDestroy(h->halo);
  HaloDestroy(h->haloFace);
  HaloDestroy(h->haloUpdate);
  if (h->dBcStates) cudaFree(h->dBcStates);
  if (h->dFlag) cudaFree(h->dFlag);
  if (h->dPartials) cudaFree(h->dPartials);
  if (h->dLinfPartials) cudaFree(h->dLinfPartials);
  if (h->dResults) cudaFree(h->dResults);
  if (h->hResults) cudaFreeHost(h->hResults);
  if (h->dStage) cudaFree(h->dStage);
  if (h->dStage2) cudaFree(h->dStage2);
  if (h->evCopied) cudaEventDestroy(h->evCopied);
  if (h->evConverted) cudaEventDestroy(h->evConverted);
  if (h->copyStream) cudaStreamDestroy(h->copyStream);
  if (h->dHaloSignal) cudaFree(h->dHaloSignal);
  if (h->evCompute) cudaEventDestroy(h->evCompute);
  if (h->evExchanged) cudaEventDestroy(h->evExchanged);
  if (h->commStream) cudaStreamDestroy(h->commStream);
  if (h->evStart) cudaEventDestroy(h->evStart);
  if (h->evStop) cudaEventDestroy(h->evStop);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const AitherEqOps *EqOpsFor(const aither_gpu *h) {
  if (h->ns == 1) return h->nt == 0 ? AitherEqOps_1_0() : AitherEqOps_1_2();
  if (h->ns == 2) return h->nt == 0 ? AitherEqOps_2_0() : AitherEqOps_2_2();
  return h->nt == 0 ? AitherEqOps_3_0() : AitherEqOps_3_2();
}

}  // namespace

#define AITHER_DEFINE_EQ_OPS(NS, NT)                                                          \
  const AitherEqOps *AitherEqOps_##NS##_##NT() {                                              \
    static const AitherEqOps ops = {PhaseBoundaryConditionsT<NS, NT>, PhaseResidualT<NS, NT>, \
                                    PhasePrepT<NS, NT>,               PhaseRelaxT<NS, NT>,    \
                                    PhaseUpdateT<NS, NT>,             StoreOldT<NS, NT>,      \
                                    InitAuxT<NS, NT>,                 OutputVarT<NS, NT>};    \
    return &ops;                                                                              \
  }
#if defined(AITHER_EQ_TU)
#if AITHER_EQ_TU == 10
AITHER_DEFINE_EQ_OPS(1, 0)
#elif AITHER_EQ_TU == 12
AITHER_DEFINE_EQ_OPS(1, 2)
#elif AITHER_EQ_TU == 20
AITHER_DEFINE_EQ_OPS(2, 0)
#elif AITHER_EQ_TU == 22
AITHER_DEFINE_EQ_OPS(2, 2)
#elif AITHER_EQ_TU == 30
AITHER_DEFINE_EQ_OPS(3, 0)
#elif AITHER_EQ_TU == 32
AITHER_DEFINE_EQ_OPS(3, 2)
#else
#error "AITHER_EQ_TU must be 10, 12, 20, 22, 30 or 32"
#endif
#elif !defined(AITHER_MAIN_TU)
AITHER_DEFINE_EQ_OPS(1, 0)
AITHER_DEFINE_EQ_OPS(1, 2)
AITHER_DEFINE_EQ_OPS(2, 0)
AITHER_DEFINE_EQ_OPS(2, 2)
AITHER_DEFINE_EQ_OPS(3, 0)
AITHER_DEFINE_EQ_OPS(3, 2)
#endif

#if !defined(AITHER_EQ_TU)  // the C ABI lives in the main translation unit only
#include "multigrid.cuh"
#include "walldist.cuh"
// =================================================================================================
extern "C" {

const char *aither_gpu_last_error(void) { return g_lastError.c_str(); }
const char *aither_gpu_version(void) { return "aither_b200 0.1 (sm_100a)"; }
const char *aither_gpu_kernel_family_name(int f) {
  return (f >= 0 && f < kNumFamilies) ? kFamilyNames[f] : "";
}
int aither_gpu_num_kernel_families(void) { return kNumFamilies; }

int aither_gpu_create(const aither_cfg *cfg, int nLocalBlocks, const aither_block_desc *blocks,
                      int nConnections, const aither_conn *conns, int rank, int nRanks,
                      void *ncclComm, int device, aither_gpu **out) {
  if (!cfg || !blocks || !out || nLocalBlocks < 1) return Fail("aither_gpu_create: bad arguments");
  std::string why;
  if (!Supported(*cfg, &why)) return Fail("aither_gpu_create: " + why);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return Fail("aither_gpu_create: no CUDA device available (there is no CPU fallback)");
  CK(cudaSetDevice(device));
  aither_gpu *h = new aither_gpu();
  h->cfg = *cfg;
  h->device = device;
  h->rank = rank;
  h->nRanks = nRanks;
  h->ns = cfg->numSpecies;
  h->nt = cfg->numTurb;
  for (int q = 0; q < cfg->numBCStates; ++q)
    h->nonreflecting = h->nonreflecting || cfg->bcStates[q].isNonreflecting != 0;
  for (int q = 0; q < cfg->numBCStates; ++q)
    h->wallLaw = h->wallLaw || (cfg->bcStates[q].isWallLaw != 0 && cfg->isViscous);
  h->neq = h->ns + 4 + h->nt;
  h->jac = cfg->isBlockMatrix ? kJacBlock
                              : (cfg->invFluxJac == AITHER_JAC_APPROX_ROE ? kJacRoe : kJacScalar);
  // scalar diagonal {flow, turbulence}, or the (ns + 4)^2 flow block + nt^2 turbulence block
  h->asz = cfg->isBlockMatrix ? (h->ns + 4) * (h->ns + 4) + h->nt * h->nt : 1 + (h->nt > 0 ? 1 : 0);
  Params &p = h->params;
  for (int s = 0; s < AITHER_MAX_SPECIES; ++s) {
    p.gas.R[s] = cfg->gasConstant[s];
    p.gas.n[s] = cfg->n[s];
    p.gas.hf[s] = cfg->hf[s];
  }
  p.kappa = cfg->kappa;
  p.theta = cfg->theta;
  p.zeta = cfg->zeta;
  p.relax = cfg->matrixRelaxation;
  p.dualTimeCFL = cfg->dualTimeCFL;
  p.dtNondim = cfg->dtNondim;
  p.isMultilevelTime = cfg->isMultilevelTime;
  p.matrixRequiresInit = cfg->matrixRequiresInit;
  p.wenoZ = cfg->recon == AITHER_RECON_WENOZ;
  p.isViscous = cfg->isViscous;
  {
    // planes ahead of the L2 prefetch in the marching kernels (0 = off; A/B switch)
    const char *pf = getenv("AITHER_B200_PREFETCH");
    p.prefetch = pf != nullptr ? std::max(0, std::min(4, atoi(pf))) : 1;

  }
  {
    const char *tv = getenv("AITHER_B200_KEEP_TIME_N");  // A/B switch: always store / read U^n
    bool nonreflecting = false;  // reads U^n in the boundary-adjacent cells
    for (int q = 0; q < cfg->numBCStates; ++q)
      nonreflecting = nonreflecting || cfg->bcStates[q].isNonreflecting != 0;
    p.timeTermsVanish = !cfg->isMultilevelTime && cfg->nonlinearIterations <= 1 &&
                        !(tv != nullptr && std::string(tv) == "1") && !nonreflecting;
  }
  p.viscRecon = cfg->viscRecon;
  p.viscCFLCoeff = cfg->viscousCFLCoeff;
  p.tr.tRef = cfg->tRef;
  for (int q = 0; q < AITHER_MAX_SPECIES; ++q) {
    p.tr.viscC1[q] = cfg->suthViscC1[q];
    p.tr.viscS[q] = cfg->suthViscS[q];
    p.tr.condC1[q] = cfg->suthCondC1[q];
    p.tr.condS[q] = cfg->suthCondS[q];
    p.tr.molarMass[q] = cfg->molarMass[q];
  }
  p.tr.schmidt = cfg->schmidt;
  p.tr.muRef = cfg->muMixRef;
  p.tr.kRef = cfg->kMixRef;
  p.tr.scaling = cfg->nondimScaling;
  p.tr.turbModel = cfg->turbModel;
  if (cfg->isViscous && !(cfg->muMixRef > 0.0 && cfg->kMixRef > 0.0 && cfg->tRef > 0.0)) {
    delete h;
    return Fail("aither_gpu_create: viscous run without transport reference values "
                "(tRef, muMixRef, kMixRef)");
  }
  GasFinalize(&p.gas);
  {
    const char *kv = getenv("AITHER_B200_KERNELS");
    h->legacyKernels = kv != nullptr && std::string(kv) == "legacy";
    // the TMA-fed sweep is inviscid-only so far; viscous runs take the register-fed march kernel
    h->tmaImplicit = !(kv != nullptr && std::string(kv) == "march") && !cfg->isViscous &&
                     cfg->numSpecies == 1;
    const char *lg = getenv("AITHER_B200_LUSGS_GRAPH");
    h->lusgsGraphs = !(lg != nullptr && std::string(lg) == "0");
    const char *ls = getenv("AITHER_B200_LUSGS_SPLIT");
    h->lusgsSplit = !(ls != nullptr && std::string(ls) == "0");
    const char *lw = getenv("AITHER_B200_LUSGS");
    h->lusgsWave = !(lw != nullptr && std::string(lw) == "planes");
    const char *fp = getenv("AITHER_B200_FUSE_PREP");
    h->fusePrep = !(fp != nullptr && std::string(fp) == "0");
  }
#define CKH(call)        \
  do {                   \
    if ((call)) {        \
      FreeAll(h);        \
      return 1;          \
    }                    \
  } while (0)
#define CKC(call)                                                      \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) {                                           \
      Fail(std::string(#call) + ": " + cudaGetErrorString(e_));        \
      FreeAll(h);                                                      \
      return 1;                                                        \
    }                                                                  \
  } while (0)
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    int prLow = 0, prHigh = 0;
    CKC(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
    CKC(cudaStreamCreateWithPriority(&h->commStream, cudaStreamNonBlocking, prHigh));
    CKC(cudaEventCreateWithFlags(&h->evCompute, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->evExchanged, cudaEventDisableTiming));
    const char *ho = getenv("AITHER_B200_HALO_OVERLAP");
    // opt-in: measured SLOWER on B200 (profiles/r02q): the sweep's 1024 tiles are 6.92 waves of
    // 148 thread blocks, and every SM lent to the pack / NCCL / unpack kernels costs an eighth wave
    h->haloOverlap = ho != nullptr && std::string(ho) == "1" && StreamWaitValueFn() != nullptr;
    CKC(cudaMalloc(&h->dHaloSignal, sizeof(unsigned int)));
    CKC(cudaMemset(h->dHaloSignal, 0, sizeof(unsigned int)));
  }
  CKC(cudaEventCreate(&h->evStart));
  CKC(cudaEventCreate(&h->evStop));
  CKC(cudaMalloc(&h->dFlag, sizeof(int)));
  CKC(cudaMemset(h->dFlag, 0, sizeof(int)));
  CKC(cudaMalloc(&h->dBcStates, sizeof(aither_bc_state) * AITHER_MAX_BC_STATES));
  CKC(cudaMemcpy(h->dBcStates, cfg->bcStates, sizeof(aither_bc_state) * AITHER_MAX_BC_STATES,
                 cudaMemcpyHostToDevice));
  if (nConnections > 0 && conns) h->conns.assign(conns, conns + nConnections);

  const int g = cfg->numGhosts;
  const int neq = h->neq;
  size_t maxCellBlocks = 0;
  h->blocks.resize(nLocalBlocks);
  for (int bb = 0; bb < nLocalBlocks; ++bb) {
    HostBlock &hb = h->blocks[bb];
    const aither_block_desc &d = blocks[bb];
    if (d.ni < 1 || d.nj < 1 || d.nk < 1) {
      Fail("aither_gpu_create: empty block");
      FreeAll(h);
      return 1;
    }
    BlockDev &b = hb.dev;
    memset(&b, 0, sizeof(b));
    b.ni = d.ni; b.nj = d.nj; b.nk = d.nk; b.g = g;
    b.parentBlock = d.parentBlock;
    hb.globalPos = d.globalPos;
    b.lp = 16;  // >= g, multiple of 16 doubles: cell i = 0 sits on a 128-byte line
    b.sj = ((b.lp + d.ni + g + 1 + 15) / 16) * 16;
    b.sk = static_cast<long long>(b.sj) * (d.nj + 2 * g + 1);
    b.fs = ((b.sk * (d.nk + 2 * g + 1) + 15) / 16) * 16;
    hb.paddedCells = static_cast<long long>(d.ni + 2 * g) * (d.nj + 2 * g) * (d.nk + 2 * g);
    // field budget (doubles per cell): state, consN, [consNm1], resid, rhs, x, xalt, [mres],
    // specRad 2, dt, diag, dinv, vol, cw 3, fA 12, center 3
    const int nFields = neq * 7 + (h->tmaImplicit ? neq : 0) + (cfg->isMultilevelTime ? neq : 0) +
                        2 + 1 + 1 + 1 + 1 + 3 + 6 +
                        12 + 3 + (cfg->isViscous ? 6 : 0) + 2 * (h->asz - 1) +
                        (h->nt > 0 ? 18 : ((cfg->isViscous && (cfg->isBlockMatrix || h->ns > 1 || h->wallLaw)) ? 9 : 0)) +
                        (h->nonreflecting ? 12 : 0);
    hb.allocBytes = static_cast<size_t>(nFields) * b.fs * sizeof(double);
    hb.nFields = nFields;
    CKC(cudaMalloc(&hb.alloc, hb.allocBytes));
    CKC(cudaMemsetAsync(hb.alloc, 0, hb.allocBytes, h->stream));
    double *cur = static_cast<double *>(hb.alloc);
    auto take = [&](int n) { double *r = cur; cur += static_cast<size_t>(n) * b.fs; return r; };
    b.state = take(neq);
    b.stateAlt = h->tmaImplicit ? take(neq) : nullptr;
    b.consN = take(neq);
    b.consNm1 = cfg->isMultilevelTime ? take(neq) : nullptr;
    b.resid = take(neq);
    b.rhs = take(neq);
    b.x = take(neq);
    b.xalt = take(neq);
    b.mres = take(neq);
    b.specRad = take(2);
    b.dt = take(1);
    b.diag = take(h->asz);
    b.dinv = take(h->asz);
    b.vol = take(1);
    for (int q = 0; q < 3; ++q) b.cw[q] = take(1);
    for (int q = 0; q < 3; ++q) b.mc[q] = take(2);
    for (int q = 0; q < 3; ++q) b.fA[q] = take(4);
    b.center = take(3);
    if (cfg->isViscous) {
      b.temperature = take(1);
      b.viscosity = take(1);
      hb.wallDistSlot = take(1);  // filled by the caller's array or by aither_gpu_compute_wall_distance
      b.wallDist = d.wallDist ? hb.wallDistSlot : nullptr;
      for (int q = 0; q < 3; ++q) b.dist[q] = take(1);
    }
    if (h->nt == 0 && cfg->isViscous && (cfg->isBlockMatrix || h->ns > 1 || h->wallLaw))
      b.velGrad = take(9);  // written by the per-cell viscous pass (RansCellKernel)
    if (h->nonreflecting) {
      b.pressGrad = take(3);
      if (h->nt == 0 && b.velGrad == nullptr) b.velGrad = take(9);
    }
    if (h->nt > 0) {
      b.eddyVisc = take(1);
      b.f1 = take(1);
      b.f2 = take(1);
      b.velGrad = take(9);
      b.tkeGrad = take(3);
      b.omegaGrad = take(3);
      if (!d.wallDist) {
        Fail("aither_gpu_create: RANS runs need the wall distance");
        FreeAll(h);
        return 1;
      }
    }

    const int NI = d.ni + 2 * g, NJ = d.nj + 2 * g, NK = d.nk + 2 * g;
    if (!d.state || !d.vol || !d.fAreaI || !d.fAreaJ || !d.fAreaK || !d.cellWidthI ||
        !d.cellWidthJ || !d.cellWidthK) {
      Fail("aither_gpu_create: block is missing a required array");
      FreeAll(h);
      return 1;
    }
    CKH(UploadAos(h, hb, d.state, NI, NJ, NK, neq, b.state, -g, -g, -g));
    CKH(UploadAos(h, hb, d.vol, NI, NJ, NK, 1, b.vol, -g, -g, -g));
    CKH(UploadAos(h, hb, d.fAreaI, NI + 1, NJ, NK, 4, b.fA[0], -g, -g, -g));
    CKH(UploadAos(h, hb, d.fAreaJ, NI, NJ + 1, NK, 4, b.fA[1], -g, -g, -g));
    CKH(UploadAos(h, hb, d.fAreaK, NI, NJ, NK + 1, 4, b.fA[2], -g, -g, -g));
    CKH(UploadAos(h, hb, d.cellWidthI, NI, NJ, NK, 1, b.cw[0], -g, -g, -g));
    CKH(UploadAos(h, hb, d.cellWidthJ, NI, NJ, NK, 1, b.cw[1], -g, -g, -g));
    CKH(UploadAos(h, hb, d.cellWidthK, NI, NJ, NK, 1, b.cw[2], -g, -g, -g));
    if (d.center) CKH(UploadAos(h, hb, d.center, NI, NJ, NK, 3, b.center, -g, -g, -g));
    if (cfg->isViscous) {
      if (!d.center) {
        Fail("aither_gpu_create: viscous runs need the cell centres");
        FreeAll(h);
        return 1;
      }
      if (d.wallDist) CKH(UploadAos(h, hb, d.wallDist, NI, NJ, NK, 1, b.wallDist, -g, -g, -g));
      {
        ScopedLaunch sl(h, kFamLayout);
        DistKernel<<<148 * 8, 256, 0, h->stream>>>(b);
      }
      EQ_DISPATCH(h, InitAuxT, h, bb);
    }
    for (int q = 0; q < 3; ++q) {
      ScopedLaunch sl(h, kFamLayout);
      MusclCoefKernel<<<148 * 8, 256, 0, h->stream>>>(b, q);
    }
    CKC(cudaGetLastError());

    // boundary surfaces
    hb.surfaces.assign(d.surfaces, d.surfaces + d.numSurfaces);
    std::vector<SurfDev> sd;
    long long off = 0;
    const int nd[3] = {d.ni, d.nj, d.nk};
    std::vector<std::vector<uint8_t>> connMask(6);
    for (int s = 0; s < 6; ++s) {
      const int d3 = s / 2, d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
      connMask[s].assign(static_cast<size_t>(nd[d1]) * nd[d2], 0);
    }
    bool anyConn = false;
    for (const auto &sf : hb.surfaces) {
      const int st = SurfaceType(sf);
      const int d3 = (st - 1) / 2, d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
      int lo[3] = {sf.imin, sf.jmin, sf.kmin}, hi[3] = {sf.imax, sf.jmax, sf.kmax};
      hb.surfFaceOffset.push_back(
          (sf.type == AITHER_BC_INTERBLOCK || sf.type == AITHER_BC_PERIODIC) ? -1 : off / g);
      if (sf.type == AITHER_BC_INTERBLOCK || sf.type == AITHER_BC_PERIODIC) {
        anyConn = true;
        for (int c2 = lo[d2]; c2 < hi[d2]; ++c2)
          for (int c1 = lo[d1]; c1 < hi[d1]; ++c1)
            connMask[st - 1][c1 + static_cast<size_t>(nd[d1]) * c2] = 1;
        continue;
      }
      SurfDev v;
      v.type = sf.type;
      v.surfType = st;
      v.tag = sf.tag;
      v.bcIndex = 0;
      const bool needsData = (sf.type == AITHER_BC_VISCOUS_WALL && cfg->isViscous) ||
                             sf.type == AITHER_BC_CHARACTERISTIC || sf.type == AITHER_BC_INLET ||
                             sf.type == AITHER_BC_SUPERSONIC_INFLOW ||
                             sf.type == AITHER_BC_STAGNATION_INLET ||
                             sf.type == AITHER_BC_PRESSURE_OUTLET;
      bool found = false;
      for (int q = 0; q < cfg->numBCStates; ++q)
        if (cfg->bcStates[q].tag == sf.tag) { v.bcIndex = q; found = true; break; }
      if (needsData && !found) {
        Fail("aither_gpu_create: boundary surface references tag " + std::to_string(sf.tag) +
             " with no boundary state");
        FreeAll(h);
        return 1;
      }
      if (sf.type != AITHER_BC_SLIP_WALL && sf.type != AITHER_BC_VISCOUS_WALL && !needsData &&
          sf.type != AITHER_BC_SUPERSONIC_OUTFLOW) {
        Fail("aither_gpu_create: unsupported boundary condition type " + std::to_string(sf.type));
        FreeAll(h);
        return 1;
      }
      for (int q = 0; q < 3; ++q) { v.lo[q] = lo[q]; v.hi[q] = hi[q]; }
      v.hi[d3] = v.lo[d3] + 1;
      v.faceOffset = off;
      off += static_cast<long long>(hi[d1] - lo[d1]) * (hi[d2] - lo[d2]) * g;
      sd.push_back(v);
    }
    {
      std::vector<EdgeSurf> es;
      long long faceBase = 0;  // SurfDev::faceOffset / g of the same surface
      for (const auto &sf : hb.surfaces) {
        EdgeSurf e;
        {
          const int st = SurfaceType(sf);
          const int d3 = (st - 1) / 2, d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
          const int lo[3] = {sf.imin, sf.jmin, sf.kmin}, hi[3] = {sf.imax, sf.jmax, sf.kmax};
          if (sf.type == AITHER_BC_INTERBLOCK || sf.type == AITHER_BC_PERIODIC) {
            e.faceBase = -1;
          } else {
            e.faceBase = static_cast<int>(faceBase);
            faceBase += static_cast<long long>(hi[d1] - lo[d1]) * (hi[d2] - lo[d2]);
          }
        }
        e.type = sf.type;
        e.surfType = SurfaceType(sf);
        e.tag = sf.tag;
        e.bcIndex = 0;
        for (int q = 0; q < cfg->numBCStates; ++q)
          if (cfg->bcStates[q].tag == sf.tag) { e.bcIndex = q; break; }
        e.lo[0] = sf.imin; e.lo[1] = sf.jmin; e.lo[2] = sf.kmin;
        e.hi[0] = sf.imax; e.hi[1] = sf.jmax; e.hi[2] = sf.kmax;
        es.push_back(e);
      }
      hb.nEdgeSurfs = static_cast<int>(es.size());
      if (!es.empty()) {
        CKC(cudaMalloc(&hb.dEdgeSurfs, sizeof(EdgeSurf) * es.size()));
        CKC(cudaMemcpy(hb.dEdgeSurfs, es.data(), sizeof(EdgeSurf) * es.size(),
                       cudaMemcpyHostToDevice));
      }
    }
    hb.nBcSurfs = static_cast<int>(sd.size());
    hb.bcThreads = off;
    if (h->nonreflecting && !sd.empty()) {
      CKC(cudaMalloc(&hb.dPatchMach, sizeof(double) * 2 * sd.size()));
      CKC(cudaMemset(hb.dPatchMach, 0, sizeof(double) * 2 * sd.size()));
    }
    {
      // wall-law walls: one record per boundary face of the block (walllaw.cuh)
      bool wallLaw = false;
      for (const auto &v : sd)
        wallLaw = wallLaw || (v.type == AITHER_BC_VISCOUS_WALL && cfg->isViscous &&
                              cfg->bcStates[v.bcIndex].isWallLaw);
      if (wallLaw && off > 0) {
        const size_t bytes = sizeof(double) * kWallVarsStride * static_cast<size_t>(off / g);
        CKC(cudaMalloc(&hb.dWallVars, bytes));
        CKC(cudaMemset(hb.dWallVars, 0, bytes));
        b.wallVars = hb.dWallVars;
      }
    }
    if (!sd.empty()) {
      CKC(cudaMalloc(&hb.dSurfs, sizeof(SurfDev) * sd.size()));
      CKC(cudaMemcpy(hb.dSurfs, sd.data(), sizeof(SurfDev) * sd.size(), cudaMemcpyHostToDevice));
    }
    if (anyConn) {
      for (int s = 0; s < 6; ++s) {
        CKC(cudaMalloc(&hb.dConnFace[s], connMask[s].size()));
        CKC(cudaMemcpy(hb.dConnFace[s], connMask[s].data(), connMask[s].size(),
                       cudaMemcpyHostToDevice));
        b.connFace[s] = hb.dConnFace[s];
      }
    }
    hb.cellBlock = dim3(32, 8, 1);
    hb.cellGrid = dim3((d.ni + 31) / 32, (d.nj + 7) / 8, d.nk);
    hb.nCellBlocks = hb.cellGrid.x * hb.cellGrid.y * hb.cellGrid.z;
    hb.cell128Grid = dim3((d.ni + 31) / 32, (d.nj + 3) / 4, d.nk);
    hb.updGrid = dim3((d.ni + 31) / 32, (d.nj + 7) / 8, (d.nk + kUpdPlanes - 1) / kUpdPlanes);
    hb.nUpdBlocks = hb.updGrid.x * hb.updGrid.y * hb.updGrid.z;
    hb.resBlock = dim3(kTI, kTJ, kTK);
    hb.resGrid = dim3((d.ni + 1 + kTI - 1) / kTI, (d.nj + 1 + kTJ - 1) / kTJ,
                      (d.nk + 1 + kTK - 1) / kTK);
    {
      // k-chunks of the marching kernels: 148 SMs x 2 resident blocks; the chunk length that
      // minimises (waves of blocks) x (planes per chunk + prologue), see PickChunk
      const int cols = ((d.ni + kMI - 1) / kMI) * ((d.nj + kMJ - 1) / kMJ);
      int chunk = PickChunk(d.nk, cols, 148 * 2, 4);
      int nChunks;
      if (const char *ev = getenv("AITHER_B200_RES_CHUNK")) chunk = std::max(1, atoi(ev));
      nChunks = (d.nk + chunk - 1) / chunk;
      hb.kChunk = chunk;
      hb.marchGrid = dim3((d.ni + kMI - 1) / kMI, (d.nj + kMJ - 1) / kMJ, nChunks);
      hb.nMarchBlocks = hb.marchGrid.x * hb.marchGrid.y * hb.marchGrid.z;
    }
    {
      // TMA-fed implicit sweep: 32 x 16 columns, chunks of ~32 planes (2 extra end planes each)
      const int cols = ((d.ni + kQI - 1) / kQI) * ((d.nj + kQJ - 1) / kQJ);
      int chunk = PickChunk(d.nk, cols, 148, 2);
      int nChunks;
      if (const char *ev = getenv("AITHER_B200_TMA_CHUNK")) chunk = std::max(1, atoi(ev));
      nChunks = (d.nk + chunk - 1) / chunk;
      hb.tmaChunk = chunk;
      hb.tmaGrid = dim3((d.ni + kQI - 1) / kQI, (d.nj + kQJ - 1) / kQJ, nChunks);
      hb.nTmaBlocks = hb.tmaGrid.x * hb.tmaGrid.y * hb.tmaGrid.z;
      std::string err;
      if (h->tmaImplicit && h->nt == 0 && h->ns == 1 &&
          (EncodeBlockMap(&hb.tmaMaps.cell, b, hb.alloc, nFields, kQPI, kQPJ, neq, &err) ||
          EncodeBlockMap(&hb.tmaMaps.faceI, b, hb.alloc, nFields, kQAI, kQJ, 4, &err) ||
          EncodeBlockMap(&hb.tmaMaps.faceJ, b, hb.alloc, nFields, kQI, kQJ + 1, 4, &err))) {
        Fail("aither_gpu_create: " + err);
        FreeAll(h);
        return 1;
      }
    }
    if (anyConn && h->tmaImplicit) {
      // tiles of the TMA sweep that hold cells a connection donates (the g layers next to a block
      // face with a connection patch; a face with any patch counts as a whole) -- computed first,
      // so that their exchange overlaps with the remaining tiles
      bool faceConn[6];
      for (int sf = 0; sf < 6; ++sf)
        faceConn[sf] = std::find(connMask[sf].begin(), connMask[sf].end(), 1) != connMask[sf].end();
      const int ext[3] = {kQI, kQJ, hb.tmaChunk};
      const int nt[3] = {static_cast<int>(hb.tmaGrid.x), static_cast<int>(hb.tmaGrid.y),
                         static_cast<int>(hb.tmaGrid.z)};
      std::vector<int> bnd, inner;
      for (int tz = 0; tz < nt[2]; ++tz)
        for (int ty = 0; ty < nt[1]; ++ty)
          for (int tx = 0; tx < nt[0]; ++tx) {
            const int t3[3] = {tx, ty, tz};
            bool touches = false;
            for (int q = 0; q < 3; ++q) {
              const int lo = t3[q] * ext[q], hi = std::min(nd[q], lo + ext[q]);  // cells [lo, hi)
              if (faceConn[2 * q] && lo < g) touches = true;
              if (faceConn[2 * q + 1] && hi > nd[q] - g) touches = true;
            }
            (touches ? bnd : inner).push_back(tx + nt[0] * (ty + nt[1] * tz));
          }
      hb.nTmaBoundaryTiles = static_cast<int>(bnd.size());
      bnd.insert(bnd.end(), inner.begin(), inner.end());
      CKC(cudaMalloc(&hb.dTmaTiles, sizeof(int) * bnd.size()));
      CKC(cudaMemcpy(hb.dTmaTiles, bnd.data(), sizeof(int) * bnd.size(), cudaMemcpyHostToDevice));
    }
    maxCellBlocks = std::max<size_t>(
        maxCellBlocks, std::max<size_t>(std::max(hb.nCellBlocks, hb.nTmaBlocks),
                                        static_cast<size_t>(hb.cell128Grid.x) * hb.cell128Grid.y *
                                            hb.cell128Grid.z));
  }
  h->partialsCap = maxCellBlocks;
  CKC(cudaMalloc(&h->dPartials, sizeof(double) * maxCellBlocks * (AITHER_MAX_SPECIES + 6)));
  CKC(cudaMalloc(&h->dLinfPartials, sizeof(LinfCand) * maxCellBlocks));
  CKH(EnsureResults(h, 64));
  // halo plan for the connections this rank takes part in
  {
    std::vector<const BlockDev *> devs;
    std::vector<int> gpos;
    for (auto &hb : h->blocks) { devs.push_back(&hb.dev); gpos.push_back(hb.globalPos); }
    const char *he = getenv("AITHER_B200_HALO_EDGES");  // A/B switch: 1 = always the full plan
    h->stateNeedsEdges = cfg->isViscous != 0 || h->nonreflecting ||
                         (he != nullptr && std::string(he) == "1");
    const char *hl = getenv("AITHER_B200_HALO_LAYERS");
    h->updateOneLayer = !(hl != nullptr && std::string(hl) == "all") &&
                        !(he != nullptr && std::string(he) == "1");
    if (HaloBuild(h->halo, h->conns, devs, gpos, neq, g, rank, nRanks, ncclComm) ||
        HaloBuild(h->haloFace, h->conns, devs, gpos, neq, g, rank, nRanks, ncclComm,
                  !(he != nullptr && std::string(he) == "1")) ||
        (h->updateOneLayer &&
         HaloBuild(h->haloUpdate, h->conns, devs, gpos, neq, 1, rank, nRanks, ncclComm, true))) {
      Fail(HaloError());
      FreeAll(h);
      return 1;
    }
  }
  CKC(cudaStreamSynchronize(h->stream));
#undef CKH
#undef CKC
  *out = h;
  return 0;
}

// ---- multigrid transfer operators (multigrid.cuh) ------------------------------
int aither_gpu_set_transfer(aither_gpu *h, int blk, const int *toCoarse, const double *volFac,
                            const double *prolong) {
  if (!h || !toCoarse || !volFac || !prolong) return Fail("null argument");
  CK(cudaSetDevice(h->device));
  if (blk < 0 || blk >= static_cast<int>(h->blocks.size())) return Fail("bad block index");
  HostBlock &hb = h->blocks[blk];
  const BlockDev &b = hb.dev;
  const size_t nc = static_cast<size_t>(b.ni) * b.nj * b.nk;
  hb.toCoarseHost.assign(toCoarse, toCoarse + 3 * nc);
  if (!hb.dToCoarse) {
    CK(cudaMalloc(&hb.dToCoarse, sizeof(int) * 3 * nc));
    CK(cudaMalloc(&hb.dVolFac, sizeof(double) * nc));
    CK(cudaMalloc(&hb.dProlong, sizeof(double) * 7 * nc));
  }
  CK(cudaMemcpy(hb.dToCoarse, toCoarse, sizeof(int) * 3 * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(hb.dVolFac, volFac, sizeof(double) * nc, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(hb.dProlong, prolong, sizeof(double) * 7 * nc, cudaMemcpyHostToDevice));
  hb.childrenFor = 0;
  return 0;
}

namespace {
int MgCheckPair(aither_gpu *fine, aither_gpu *coarse) {
  if (!fine || !coarse) return Fail("null handle");
  if (fine->device != coarse->device) return Fail("multigrid levels must live on one device");
  if (fine->blocks.size() != coarse->blocks.size() || fine->neq != coarse->neq)
    return Fail("multigrid levels must have the same blocks and equations");
  for (auto &hb : fine->blocks)
    if (!hb.dToCoarse) return Fail("aither_gpu_set_transfer has not been called for the fine level");
  return 0;
}
// children of every coarse cell in the reference's visiting order (fine k, j, i ascending)
int MgChildren(HostBlock &f, const HostBlock &c) {
  const long long ncc = static_cast<long long>(c.dev.ni) * c.dev.nj * c.dev.nk;
  if (f.dChildren && f.childrenFor == ncc) return 0;
  std::vector<int> ch(static_cast<size_t>(ncc) * kMaxChildren, -1);
  std::vector<int> cnt(static_cast<size_t>(ncc), 0);
  const long long nfc = static_cast<long long>(f.dev.ni) * f.dev.nj * f.dev.nk;
  for (long long pf = 0; pf < nfc; ++pf) {
    const int *ci = &f.toCoarseHost[3 * pf];
    if (ci[0] < 0 || ci[0] >= c.dev.ni || ci[1] < 0 || ci[1] >= c.dev.nj || ci[2] < 0 ||
        ci[2] >= c.dev.nk)
      return Fail("multigrid transfer map points outside the coarse block");
    const long long pc = ci[0] + static_cast<long long>(c.dev.ni) * (ci[1] + static_cast<long long>(c.dev.nj) * ci[2]);
    if (cnt[pc] >= kMaxChildren) return Fail("a coarse cell has more than 8 fine cells");
    ch[pc * kMaxChildren + cnt[pc]++] = static_cast<int>(pf);
  }
  if (f.dChildren) cudaFree(f.dChildren);
  f.dChildren = nullptr;
  CK(cudaMalloc(&f.dChildren, sizeof(int) * ch.size()));
  CK(cudaMemcpy(f.dChildren, ch.data(), sizeof(int) * ch.size(), cudaMemcpyHostToDevice));
  f.childrenFor = ncc;
  return 0;
}
}  // namespace

int aither_gpu_mg_restrict(aither_gpu *fine, aither_gpu *coarse, int mm, double cfl) {
  // gridLevel::Restriction (ref: src/gridLevel.cpp:538-588)
  if (MgCheckPair(fine, coarse)) return 1;
  CK(cudaSetDevice(fine->device));
  CK(cudaStreamSynchronize(fine->stream));
  const int neq = fine->neq;
  const dim3 blk(32, 4, 1);
  for (size_t bb = 0; bb < fine->blocks.size(); ++bb) {
    HostBlock &f = fine->blocks[bb];
    HostBlock &c = coarse->blocks[bb];
    if (MgChildren(f, c)) return 1;
    const dim3 grid((c.dev.ni + 31) / 32, (c.dev.nj + 3) / 4, c.dev.nk);
    RestrictKernel<<<grid, blk, 0, coarse->stream>>>(f.dev, c.dev, f.dChildren, f.dVolFac,
                                                     f.dev.state, c.dev.state, neq, false);
  }
  CK(cudaGetLastError());
  coarse->stateMovedSinceStore = true;
  if (mm == 0 && aither_gpu_store_old_solution(coarse, 1)) return 1;
  // The reference's residual loops ADD the spectral radii / flux Jacobians to the main diagonal,
  // which is only zeroed by ResetDiagonal at the end of the iteration (src/mgSolution.cpp:262-265):
  // a level that is restricted to more than once per iteration (W cycle) keeps the inverted-
  // diagonal terms of its previous visit underneath the new ones. The residual kernels here
  // assign the diagonal, so the previous one is saved and added back.
  const long long nDiag = static_cast<long long>(coarse->asz);
  for (auto &c : coarse->blocks) {
    const size_t bytes = sizeof(double) * c.dev.fs * nDiag;
    if (!c.dSavedDiag) CK(cudaMalloc(&c.dSavedDiag, bytes));
    CK(cudaMemcpyAsync(c.dSavedDiag, c.dev.diag, bytes, cudaMemcpyDeviceToDevice, coarse->stream));
  }
  if (aither_gpu_get_boundary_conditions(coarse) || aither_gpu_calc_residual(coarse)) return 1;
  for (auto &c : coarse->blocks)
    AxpyFieldKernel<<<148 * 4, 256, 0, coarse->stream>>>(c.dev.diag, c.dSavedDiag, 1.0,
                                                       static_cast<long long>(c.dev.fs) * nDiag);
  CK(cudaGetLastError());
  if (aither_gpu_calc_time_step(coarse, cfl) || aither_gpu_invert_diagonal(coarse) ||
      aither_gpu_initialize_matrix_update(coarse))  // right-hand side b; x is replaced below
    return 1;
  // linearSolver::Restriction: volume-weighted update (ghosts zero), then its ghost swap
  for (size_t bb = 0; bb < fine->blocks.size(); ++bb) {
    HostBlock &f = fine->blocks[bb];
    HostBlock &c = coarse->blocks[bb];
    CK(cudaMemsetAsync(c.dev.x, 0, sizeof(double) * c.dev.fs * neq, coarse->stream));
    const dim3 grid((c.dev.ni + 31) / 32, (c.dev.nj + 3) / 4, c.dev.nk);
    RestrictKernel<<<grid, blk, 0, coarse->stream>>>(f.dev, c.dev, f.dChildren, f.dVolFac, f.dev.x,
                                                     c.dev.x, neq, false);
  }
  CK(cudaGetLastError());
  // A x - b of the coarse level with the restricted update (no sweeps: swap + matrix residual)
  double unused = 0.0;
  if (aither_gpu_relax(coarse, 0, &unused)) return 1;
  for (size_t bb = 0; bb < fine->blocks.size(); ++bb) {
    HostBlock &f = fine->blocks[bb];
    HostBlock &c = coarse->blocks[bb];
    const dim3 grid((c.dev.ni + 31) / 32, (c.dev.nj + 3) / 4, c.dev.nk);
    ForcingKernel<<<grid, blk, 0, coarse->stream>>>(f.dev, c.dev, f.dChildren, neq);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(coarse->stream));
  return 0;
}

int aither_gpu_mg_save_update(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  for (auto &hb : h->blocks) {
    const size_t bytes = sizeof(double) * hb.dev.fs * h->neq;
    if (!hb.dSavedX) CK(cudaMalloc(&hb.dSavedX, bytes));
    CK(cudaMemcpyAsync(hb.dSavedX, hb.dev.x, bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  return 0;
}

int aither_gpu_mg_subtract_saved(aither_gpu *h) {
  // linearSolver::SubtractFromUpdate (ref: src/linearSolver.cpp:195-201)
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  for (auto &hb : h->blocks) {
    if (!hb.dSavedX) return Fail("aither_gpu_mg_save_update has not been called");
    AxpyFieldKernel<<<148 * 4, 256, 0, h->stream>>>(hb.dev.x, hb.dSavedX, -1.0,
                                                   static_cast<long long>(hb.dev.fs) * h->neq);
  }
  CK(cudaGetLastError());
  return 0;
}

int aither_gpu_mg_prolong(aither_gpu *coarse, aither_gpu *fine) {
  // gridLevel::Prolongation (ref: src/gridLevel.cpp:594-611)
  if (MgCheckPair(fine, coarse)) return 1;
  CK(cudaSetDevice(fine->device));
  CK(cudaStreamSynchronize(coarse->stream));
  const int neq = fine->neq;
  for (size_t bb = 0; bb < fine->blocks.size(); ++bb) {
    HostBlock &f = fine->blocks[bb];
    HostBlock &c = coarse->blocks[bb];
    const long long nn = static_cast<long long>(c.dev.ni + 1) * (c.dev.nj + 1) * (c.dev.nk + 1);
    if (!c.dNodes) CK(cudaMalloc(&c.dNodes, sizeof(double) * nn * neq));
    const dim3 blk(32, 4, 1);
    const dim3 gridN((c.dev.ni + 1 + 31) / 32, (c.dev.nj + 1 + 3) / 4, c.dev.nk + 1);
    NodeKernel<<<gridN, blk, 0, fine->stream>>>(c.dev, c.dev.x, c.dNodes, neq);
    const dim3 gridF((f.dev.ni + 31) / 32, (f.dev.nj + 3) / 4, f.dev.nk);
    ProlongKernel<<<gridF, blk, 0, fine->stream>>>(f.dev, c.dev, f.dToCoarse, f.dProlong, c.dNodes,
                                                   f.dev.x, neq);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(fine->stream));
  return 0;
}

int aither_gpu_destroy(aither_gpu *h) {
  if (h) DrainProfile(h);
  FreeAll(h);
  return 0;
}

int aither_gpu_store_old_solution(aither_gpu *h, int iter) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  const int copyNm1 = (h->cfg.isMultilevelTime && iter == 0) ? 1 : 0;
  if (h->params.timeTermsVanish) {
    // U^n is not needed by the iteration; it is materialised only if someone asks for it
    // (aither_gpu_download_field(AITHER_FIELD_CONS_N)) before the state moves on
    h->consNStale = true;
    h->stateMovedSinceStore = false;
    return 0;
  }
  EQ_DISPATCH(h, StoreOldT, h, copyNm1);
  CK(cudaGetLastError());
  return 0;
}

int aither_gpu_iterate(aither_gpu *h, double cfl, int mm, double *residL2, aither_linf *linf,
                       double *matrixResid) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (IterateAsync(h, cfl, 0, mm)) return 1;
  if (FetchResults(h, 1)) return 1;
  const IterResult &r = h->hResults[0];
  if (residL2)
    for (int e = 0; e < h->neq; ++e) residL2[e] += r.l2[e];
  if (linf && r.linf > linf->linf) {
    linf->linf = r.linf;
    linf->block = r.linfBlock;
    linf->i = r.linfI;
    linf->j = r.linfJ;
    linf->k = r.linfK;
    linf->eqn = r.linfEqn;
  }
  if (matrixResid) *matrixResid = r.matrixSumSq / static_cast<double>(TotalPaddedSize(h));
  return 0;
}

int aither_gpu_run(aither_gpu *h, int nIter, double cflStart, double cflStep, double cflMax,
                   double *hist) {
  if (!h) return Fail("null handle");
  if (nIter < 1) return 0;
  CK(cudaSetDevice(h->device));
  const int nl = std::max(1, h->cfg.nonlinearIterations);
  if (EnsureResults(h, nIter * nl)) return 1;
  for (int n = 0; n < nIter; ++n) {
    const double cfl = std::min(cflStart + n * cflStep, cflMax);  // ref: src/input.cpp:647
    if (aither_gpu_store_old_solution(h, n)) return 1;
    for (int mm = 0; mm < nl; ++mm)
      if (IterateAsync(h, cfl, n * nl + mm, mm)) return 1;
  }
  nIter *= nl;
  if (FetchResults(h, nIter)) return 1;
  if (hist) {
    const double tot = static_cast<double>(TotalPaddedSize(h));
    for (int n = 0; n < nIter; ++n) {
      for (int e = 0; e < h->neq; ++e) hist[n * (h->neq + 1) + e] = h->hResults[n].l2[e];
      hist[n * (h->neq + 1) + h->neq] = h->hResults[n].matrixSumSq / tot;
    }
  }
  return 0;
}

int aither_gpu_get_boundary_conditions(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (PhaseBoundaryConditions(h)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int aither_gpu_calc_residual(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (PhaseResidual(h)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int aither_gpu_calc_time_step(aither_gpu *h, double cfl) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (PhasePrep(h, cfl, kPrepDt)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int aither_gpu_invert_diagonal(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (PhasePrep(h, 0.0, kPrepDiag)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return CheckFlag(h);
}
int aither_gpu_initialize_matrix_update(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (PhasePrep(h, 0.0, kPrepInit)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int aither_gpu_relax(aither_gpu *h, int sweeps, double *matrixResid) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  h->keepMatrixResid = true;
  if (ZeroResult(h, 0)) return 1;
  if (PhaseRelax(h, sweeps, 0)) return 1;
  if (FetchResults(h, 1)) return 1;
  h->keepMatrixResid = false;
  if (matrixResid) *matrixResid = h->hResults[0].matrixSumSq / static_cast<double>(TotalPaddedSize(h));
  return 0;
}
int aither_gpu_update_blocks(aither_gpu *h, int mm, double *residL2, aither_linf *linf) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (ZeroResult(h, 0)) return 1;
  if (PhaseUpdate(h, 0, mm)) return 1;
  if (FetchResults(h, 1)) return 1;
  const IterResult &r = h->hResults[0];
  if (residL2)
    for (int e = 0; e < h->neq; ++e) residL2[e] += r.l2[e];
  if (linf && r.linf > linf->linf) {
    linf->linf = r.linf;
    linf->block = r.linfBlock;
    linf->i = r.linfI;
    linf->j = r.linfJ;
    linf->k = r.linfK;
    linf->eqn = r.linfEqn;
  }
  return 0;
}
int aither_gpu_reset_diagonal(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  for (auto &hb : h->blocks) {
    CK(cudaMemsetAsync(hb.dev.diag, 0, sizeof(double) * hb.dev.fs * h->asz, h->stream));
  }
  return 0;
}

static int FieldInfo(aither_gpu *h, int blk, int field, const double **ptr, int *nc, bool *padded) {
  if (blk < 0 || blk >= static_cast<int>(h->blocks.size())) return Fail("bad block index");
  const BlockDev &b = h->blocks[blk].dev;
  switch (field) {
    case AITHER_FIELD_STATE: *ptr = b.state; *nc = h->neq; *padded = true; break;
    case AITHER_FIELD_RESIDUAL: *ptr = b.resid; *nc = h->neq; *padded = false; break;
    case AITHER_FIELD_SPEC_RADIUS: *ptr = b.specRad; *nc = 2; *padded = false; break;
    case AITHER_FIELD_DT: *ptr = b.dt; *nc = 1; *padded = false; break;
    case AITHER_FIELD_DIAG: *ptr = b.diag; *nc = h->asz; *padded = false; break;
    case AITHER_FIELD_DIAG_INV: *ptr = b.dinv; *nc = h->asz; *padded = false; break;
    case AITHER_FIELD_UPDATE: *ptr = b.x; *nc = h->neq; *padded = true; break;
    case AITHER_FIELD_CONS_N: *ptr = b.consN; *nc = h->neq; *padded = false; break;
    case AITHER_FIELD_MATRIX_RESID: *ptr = b.mres; *nc = h->neq; *padded = false; break;
    case AITHER_FIELD_CONS_NM1:
      if (!b.consNm1) return Fail("consNm1 is only stored for bdf2");
      *ptr = b.consNm1; *nc = h->neq; *padded = false; break;
    case AITHER_FIELD_TEMPERATURE:
      if (!b.temperature) return Fail("temperature is only stored for viscous runs");
      *ptr = b.temperature; *nc = 1; *padded = true; break;
    case AITHER_FIELD_VISCOSITY:
      if (!b.viscosity) return Fail("viscosity is only stored for viscous runs");
      *ptr = b.viscosity; *nc = 1; *padded = true; break;
    case AITHER_FIELD_WALL_DIST:
      if (!b.wallDist) return Fail("this run has no wall distance");
      *ptr = b.wallDist; *nc = 1; *padded = true; break;
    case AITHER_FIELD_PRESSURE_GRAD:
      if (!b.pressGrad) return Fail("the pressure gradient is only kept for runs with non-reflecting BCs");
      *ptr = b.pressGrad; *nc = 3; *padded = false; break;
    case AITHER_FIELD_VELOCITY_GRAD:
      if (!b.velGrad) return Fail("the velocity gradient is not kept by this configuration");
      *ptr = b.velGrad; *nc = 9; *padded = true; break;
    case AITHER_FIELD_EDDY_VISCOSITY: case AITHER_FIELD_F1: case AITHER_FIELD_F2:
    case AITHER_FIELD_TKE_GRAD: case AITHER_FIELD_OMEGA_GRAD:
      if (!b.eddyVisc) return Fail("turbulence fields are only stored for RANS runs");
      *padded = field != AITHER_FIELD_TKE_GRAD && field != AITHER_FIELD_OMEGA_GRAD;
      *nc = field == AITHER_FIELD_VELOCITY_GRAD ? 9 : (*padded ? 1 : 3);
      *ptr = field == AITHER_FIELD_EDDY_VISCOSITY ? b.eddyVisc
             : field == AITHER_FIELD_F1 ? b.f1
             : field == AITHER_FIELD_F2 ? b.f2
             : field == AITHER_FIELD_VELOCITY_GRAD ? b.velGrad
             : field == AITHER_FIELD_TKE_GRAD ? b.tkeGrad : b.omegaGrad;
      break;
    default: return Fail("unknown or unavailable field id " + std::to_string(field));
  }
  return 0;
}

long long aither_gpu_field_size(aither_gpu *h, int blk, int field) {
  if (!h) return -1;
  const double *ptr; int nc; bool padded;
  if (FieldInfo(h, blk, field, &ptr, &nc, &padded)) return -1;
  const BlockDev &b = h->blocks[blk].dev;
  const int g = padded ? b.g : 0;
  return static_cast<long long>(b.ni + 2 * g) * (b.nj + 2 * g) * (b.nk + 2 * g) * nc;
}

int aither_gpu_download_field(aither_gpu *h, int blk, int field, double *dst) {
  if (!h || !dst) return Fail("null argument");
  CK(cudaSetDevice(h->device));
  if (field == AITHER_FIELD_CONS_N && h->consNStale) {
    if (h->stateMovedSinceStore)
      return Fail("U^n was not stored: with one nonlinear iteration per step of a single-level "
                  "scheme the iteration does not need it (set AITHER_B200_KEEP_TIME_N=1 to keep it)");
    EQ_DISPATCH(h, StoreOldT, h, 0);
    CK(cudaGetLastError());
    h->consNStale = false;
  }
  const double *ptr; int nc; bool padded;
  if (FieldInfo(h, blk, field, &ptr, &nc, &padded)) return 1;
  HostBlock &hb = h->blocks[blk];
  const BlockDev &b = hb.dev;
  if (field == AITHER_FIELD_STATE && hb.ghostsInAlt && b.stateAlt) {
    // the ghost cells of the last fill sit in the buffer the fused update left behind
    const long long n = static_cast<long long>(b.ni + 2 * b.g) * (b.nj + 2 * b.g) * (b.nk + 2 * b.g);
    GhostShellCopyKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, h->stream>>>(
        b, b.stateAlt, b.state, h->neq);
    CK(cudaGetLastError());
    hb.ghostsInAlt = false;
  }
  const int g = padded ? b.g : 0;
  return DownloadAos(h, hb, dst, b.ni + 2 * g, b.nj + 2 * g, b.nk + 2 * g, nc, ptr, -g, -g, -g);
}
int aither_gpu_download_state(aither_gpu *h, int blk, double *stateAoS) {
  return aither_gpu_download_field(h, blk, AITHER_FIELD_STATE, stateAoS);
}
int aither_gpu_download_output(aither_gpu *h, int blk, int var, int species, double scale,
                               double *dst) {
  // ref: WriteFunFile, src/output.cpp:209-437
  if (!h || !dst) return Fail("null argument");
  CK(cudaSetDevice(h->device));
  if (blk < 0 || blk >= static_cast<int>(h->blocks.size())) return Fail("bad block index");
  if (var < 0 || var >= AITHER_OUT_NUM_VARS) return Fail("unknown output variable " + std::to_string(var));
  const BlockDev &b = h->blocks[blk].dev;
  if ((var == AITHER_OUT_VISCOSITY || var == AITHER_OUT_VISCOSITY_RATIO) && !b.viscosity)
    return Fail("viscosity is only kept for viscous runs");
  if ((var == AITHER_OUT_TURBULENT_VISCOSITY || var == AITHER_OUT_F1 || var == AITHER_OUT_F2 ||
       var == AITHER_OUT_TKE || var == AITHER_OUT_SDR) && h->nt == 0)
    return Fail("turbulence variables are only kept for RANS runs");
  if (var == AITHER_OUT_WALL_DISTANCE && !b.wallDist) return Fail("this run has no wall distance");
  if (var == AITHER_OUT_MASS_FRACTION && (species < 0 || species >= h->ns))
    return Fail("species index out of range");
  const size_t n = static_cast<size_t>(b.ni) * b.nj * b.nk;
  if (EnsureStage(h, n * sizeof(double))) return 1;
  if (EQ_DISPATCH(h, OutputVarT, h, blk, var, species, scale, h->dStage)) return 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, h->dStage, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int aither_gpu_compute_wall_distance(aither_gpu *h, const double *wallFaceCenters, long long n) {
  // ref: src/main.cpp:144,191-201, src/procBlock.cpp:6030-6107, src/kdtree.cpp:211-225
  if (!h) return Fail("null handle");
  if (n < 0 || (n > 0 && !wallFaceCenters)) return Fail("aither_gpu_compute_wall_distance: bad point list");
  if (n == 0 || !h->cfg.isViscous) return 0;  // the reference skips the search without viscous walls
  CK(cudaSetDevice(h->device));
  double *dPts = nullptr;
  CK(cudaMalloc(&dPts, sizeof(double) * 3 * static_cast<size_t>(n)));
  CK(cudaMemcpyAsync(dPts, wallFaceCenters, sizeof(double) * 3 * static_cast<size_t>(n),
                     cudaMemcpyHostToDevice, h->stream));
  for (auto &hb : h->blocks) {
    BlockDev &b = hb.dev;
    if (!hb.wallDistSlot) { cudaFree(dPts); return Fail("this handle keeps no wall-distance field"); }
    if (!b.wallDist) {
      // no array at create: the edge ghost cells, which the search does not touch, hold zero
      b.wallDist = hb.wallDistSlot;
      CK(cudaMemsetAsync(b.wallDist, 0, sizeof(double) * b.fs, h->stream));
    }
    const long long nCells = static_cast<long long>(b.ni) * b.nj * b.nk;
    ScopedLaunch sl(h, kFamLayout);
    WallDistKernel<<<static_cast<unsigned>((nCells + 255) / 256), 256, 0, h->stream>>>(b, dPts, n);
    for (const aither_surface &sf : hb.surfaces) {
      const int st = SurfaceType(sf);
      const int d3 = (st - 1) / 2, d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
      const int lo[3] = {sf.imin, sf.jmin, sf.kmin}, hi[3] = {sf.imax, sf.jmax, sf.kmax};
      const int n1 = hi[d1] - lo[d1], n2 = hi[d2] - lo[d2];
      if (n1 <= 0 || n2 <= 0) continue;
      const int total = n1 * n2 * b.g;
      WallDistGhostKernel<<<(total + 127) / 128, 128, 0, h->stream>>>(
          b, d3, st % 2, sf.type == AITHER_BC_VISCOUS_WALL ? 1 : 0, lo[d1], n1, lo[d2], n2);
    }
  }
  CK(cudaGetLastError());
  // ghost cells across connections take the neighbour block's distance (mgSolution::SwapWallDist,
  // src/main.cpp:202, src/gridLevel.cpp:261-281)
  if (Exchange(h, kHaloWallDist)) { cudaFree(dPts); return 1; }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaFree(dPts));
  return 0;
}
int aither_gpu_download_wall_data(aither_gpu *h, int blk, int surface, double *dst) {
  // ref: include/wallData.hpp:40-57, src/procBlock.cpp:6287-6290 (records of wall-law walls)
  if (!h || !dst) return Fail("null argument");
  CK(cudaSetDevice(h->device));
  if (blk < 0 || blk >= static_cast<int>(h->blocks.size())) return Fail("bad block index");
  const HostBlock &hb = h->blocks[blk];
  if (surface < 0 || surface >= static_cast<int>(hb.surfaces.size())) return Fail("bad surface index");
  const aither_surface &sf = hb.surfaces[surface];
  bool wallLaw = false;
  if (sf.type == AITHER_BC_VISCOUS_WALL && h->cfg.isViscous)
    for (int q = 0; q < h->cfg.numBCStates; ++q)
      if (h->cfg.bcStates[q].tag == sf.tag) { wallLaw = h->cfg.bcStates[q].isWallLaw != 0; break; }
  if (!wallLaw || !hb.dWallVars || hb.surfFaceOffset[surface] < 0)
    return Fail("wall data is kept for viscous walls with the wall law only (surface " +
                std::to_string(surface) + " is not one)");
  const int st = SurfaceType(sf);
  const int d3 = (st - 1) / 2, d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
  const int lo[3] = {sf.imin, sf.jmin, sf.kmin}, hi[3] = {sf.imax, sf.jmax, sf.kmax};
  const int n1 = hi[d1] - lo[d1], n2 = hi[d2] - lo[d2];
  std::vector<double> raw(static_cast<size_t>(n1) * n2 * kWallVarsStride);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(raw.data(), hb.dWallVars + hb.surfFaceOffset[surface] * kWallVarsStride,
                sizeof(double) * raw.size(), cudaMemcpyDeviceToHost));
  // device order: direction 1 fastest (i-surface: j, k; j-surface: k, i; k-surface: i, j);
  // reference order: i fastest, then j, then k
  int ext[3] = {1, 1, 1};
  ext[d1] = n1;
  ext[d2] = n2;
  for (int a2 = 0; a2 < n2; ++a2)
    for (int a1 = 0; a1 < n1; ++a1) {
      int c[3] = {0, 0, 0};
      c[d1] = a1;
      c[d2] = a2;
      const double *w = &raw[(static_cast<size_t>(a2) * n1 + a1) * kWallVarsStride];
      double *o = dst + (static_cast<size_t>(c[2]) * ext[1] * ext[0] + static_cast<size_t>(c[1]) * ext[0] + c[0]) *
                            AITHER_WALL_VARS;
      o[0] = w[kWvYplus];
      o[1] = w[kWvTau]; o[2] = w[kWvTau + 1]; o[3] = w[kWvTau + 2];
      o[4] = w[kWvHeatFlux];
      o[5] = w[kWvT];
      o[6] = w[kWvMut];
      o[7] = w[kWvMu];
      o[8] = w[kWvRho];
      o[9] = w[kWvUtau];
      o[10] = w[kWvTke];
      o[11] = w[kWvSdr];
    }
  return 0;
}
int aither_gpu_upload_state(aither_gpu *h, int blk, const double *stateAoS) {
  if (!h || !stateAoS) return Fail("null argument");
  CK(cudaSetDevice(h->device));
  if (blk < 0 || blk >= static_cast<int>(h->blocks.size())) return Fail("bad block index");
  const HostBlock &hb = h->blocks[blk];
  const BlockDev &b = hb.dev;
  const int g = b.g;
  if (UploadAos(h, hb, stateAoS, b.ni + 2 * g, b.nj + 2 * g, b.nk + 2 * g, h->neq, b.state, -g, -g,
                -g))
    return 1;
  h->blocks[blk].ghostsInAlt = false;
  h->stateMovedSinceStore = true;  // a U^n that was not materialised can no longer be
  if (h->nt > 0) {  // the wall omega BC reads the stored viscosity: make it the new state's
    EQ_DISPATCH(h, InitAuxT, h, blk);
    CK(cudaGetLastError());
  }
  return 0;
}

static int UploadStateAsync(aither_gpu *h, int blk, const double *stateAoS, bool interior) {
  if (!h || !stateAoS) return Fail("null argument");
  CK(cudaSetDevice(h->device));
  if (blk < 0 || blk >= static_cast<int>(h->blocks.size())) return Fail("bad block index");
  if (h->pendingBlk >= 0) return Fail("an asynchronous upload is already pending: commit it first");
  const BlockDev &b = h->blocks[blk].dev;
  const int gg = interior ? 0 : b.g;
  const size_t n = static_cast<size_t>(b.ni + 2 * gg) * (b.nj + 2 * gg) * (b.nk + 2 * gg) * h->neq;
  if (!h->copyStream) {
    CK(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->evCopied, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->evConverted, cudaEventDisableTiming));
    CK(cudaEventRecord(h->evConverted, h->stream));
  }
  if (n * sizeof(double) > h->stage2Bytes) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->dStage2) cudaFree(h->dStage2);
    h->dStage2 = nullptr;
    h->stage2Bytes = 0;
    CK(cudaMalloc(&h->dStage2, n * sizeof(double)));
    h->stage2Bytes = n * sizeof(double);
  }
  // the previous conversion must have finished reading the staging buffer
  CK(cudaStreamWaitEvent(h->copyStream, h->evConverted, 0));
  CK(cudaMemcpyAsync(h->dStage2, stateAoS, n * sizeof(double), cudaMemcpyHostToDevice,
                     h->copyStream));
  CK(cudaEventRecord(h->evCopied, h->copyStream));
  h->pendingBlk = blk;
  h->pendingInterior = interior;
  return 0;
}
int aither_gpu_upload_state_async(aither_gpu *h, int blk, const double *stateAoS) {
  return UploadStateAsync(h, blk, stateAoS, false);
}
int aither_gpu_upload_interior_async(aither_gpu *h, int blk, const double *interiorAoS) {
  return UploadStateAsync(h, blk, interiorAoS, true);
}
int aither_gpu_upload_state_commit(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  if (h->pendingBlk < 0) return Fail("no asynchronous upload is pending");
  const int blk = h->pendingBlk;
  h->pendingBlk = -1;
  const HostBlock &hb = h->blocks[blk];
  const BlockDev &b = hb.dev;
  // physical cells only: the ghost cells are filled at the start of every iteration anyway
  const int g = h->pendingInterior ? 0 : b.g, SI = b.ni + 2 * g, SJ = b.nj + 2 * g, SK = b.nk + 2 * g;
  CK(cudaStreamWaitEvent(h->stream, h->evCopied, 0));
  const long long cells = static_cast<long long>(SI) * SJ * SK;
  const int grid = static_cast<int>(std::min<long long>((cells + 255) / 256, 148 * 16));
  {
    ScopedLaunch sl(h, kFamLayout);
    const int off = b.g - g;  // rows / planes the source lacks in front (0 with ghost cells)
    AosToSoaKernel<<<grid, 256, 0, h->stream>>>(h->dStage2, SI, SJ, SK, h->neq, b.state, b.fs,
                                                -g + b.lp, off, off, b.sj, b.sk);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->evConverted, h->stream));
  h->blocks[blk].ghostsInAlt = false;
  h->stateMovedSinceStore = true;
  if (h->nt > 0) {
    EQ_DISPATCH(h, InitAuxT, h, blk);
    CK(cudaGetLastError());
  }
  return 0;
}

int aither_gpu_alloc_host(long long bytes, void **out) {
  if (!out || bytes <= 0) return Fail("aither_gpu_alloc_host: bad arguments");
  CK(cudaMallocHost(out, static_cast<size_t>(bytes)));
  return 0;
}
int aither_gpu_free_host(void *p) {
  if (p) CK(cudaFreeHost(p));
  return 0;
}

int aither_gpu_comm_unique_id(char id[128]) {
  if (!id) return Fail("null argument");
  NcclApi *api = Nccl();
  if (!api) return Fail(HaloError());
  NcclUniqueId u;
  const int rc = api->GetUniqueId(&u);
  if (rc != 0) return Fail(std::string("ncclGetUniqueId: ") + api->GetErrorString(rc));
  memcpy(id, u.internal, 128);
  return 0;
}
int aither_gpu_comm_create(const char id[128], int rank, int nRanks, int device, void **comm) {
  if (!id || !comm) return Fail("null argument");
  NcclApi *api = Nccl();
  if (!api) return Fail(HaloError());
  CK(cudaSetDevice(device));
  NcclUniqueId u;
  memcpy(u.internal, id, 128);
  const int rc = api->CommInitRank(comm, nRanks, u, rank);
  if (rc != 0) return Fail(std::string("ncclCommInitRank: ") + api->GetErrorString(rc));
  return 0;
}
int aither_gpu_comm_destroy(void *comm) {
  if (!comm) return 0;
  NcclApi *api = Nccl();
  if (!api) return Fail(HaloError());
  const int rc = api->CommDestroy(comm);
  if (rc != 0) return Fail(std::string("ncclCommDestroy: ") + api->GetErrorString(rc));
  return 0;
}
int aither_gpu_halo_p2p_export(aither_gpu *h, void *handle) {
  if (!h || !handle) return Fail("null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == AITHER_P2P_HANDLE_BYTES, "IPC handle size");
  CK(cudaSetDevice(h->device));
  HaloPlan *plans[3] = {&h->halo, &h->haloFace, &h->haloUpdate};
  if (!h->p2pArena) {
    h->p2pBytes = P2POffset(plans, 3, h->rank, -1, 0, 0);
    CK(cudaMalloc(&h->p2pArena, h->p2pBytes));
    CK(cudaMemset(h->p2pArena, 0, h->p2pBytes));
  }
  cudaIpcMemHandle_t ipc;
  CK(cudaIpcGetMemHandle(&ipc, h->p2pArena));
  memcpy(handle, &ipc, sizeof(ipc));
  return 0;
}
int aither_gpu_halo_p2p_import(aither_gpu *h, const void *handles) {
  if (!h || !handles) return Fail("null argument");
  if (!h->p2pArena) return Fail("aither_gpu_halo_p2p_import: call aither_gpu_halo_p2p_export first");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<unsigned char *> arena(h->nRanks, nullptr);
  h->p2pPeers.assign(h->nRanks, nullptr);
  for (int r = 0; r < h->nRanks; ++r) {
    if (r == h->rank) { arena[r] = h->p2pArena; continue; }
    // only the ranks this one exchanges with are mapped
    bool partner = false;
    for (const aither_conn &c : h->conns)
      partner = partner || (c.rank[0] == h->rank && c.rank[1] == r) || (c.rank[1] == h->rank && c.rank[0] == r);
    if (!partner) continue;
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, static_cast<const unsigned char *>(handles) + static_cast<size_t>(r) * sizeof(ipc), sizeof(ipc));
    void *ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return Fail(std::string("aither_gpu_halo_p2p_import: cudaIpcOpenMemHandle failed for rank ") +
                  std::to_string(r) + ": " + cudaGetErrorString(e));
    h->p2pPeers[r] = static_cast<unsigned char *>(ptr);
    arena[r] = h->p2pPeers[r];
  }
  HaloPlan *plans[3] = {&h->halo, &h->haloFace, &h->haloUpdate};
  if (HaloP2PEnable(plans, 3, arena)) return Fail(HaloError());
  return 0;
}
int aither_gpu_halo_info(aither_gpu *h, int *levels, long long *remoteCells) {
  if (!h) return Fail("null handle");
  if (levels) *levels = static_cast<int>(h->halo.levels.size());  // the reference-order plan
  if (remoteCells) *remoteCells = h->halo.bytesPerExchangeRemote;
  return 0;
}

int aither_gpu_synchronize(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
int aither_gpu_timer_start(aither_gpu *h) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->evStart, h->stream));
  return 0;
}
int aither_gpu_timer_stop(aither_gpu *h, float *ms) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->evStop, h->stream));
  CK(cudaEventSynchronize(h->evStop));
  if (ms) CK(cudaEventElapsedTime(ms, h->evStart, h->evStop));
  return 0;
}
long long aither_gpu_launch_count(aither_gpu *h) { return h ? h->launches : -1; }
int aither_gpu_profile_enable(aither_gpu *h, int enable) {
  if (!h) return Fail("null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  DrainProfile(h);
  h->profile = enable != 0;
  for (int f = 0; f < kNumFamilies; ++f) {
    h->famMs[f] = 0.0;
    h->famLaunches[f] = 0;
  }
  return 0;
}
int aither_gpu_profile_get(aither_gpu *h, int family, double *ms, long long *launches) {
  if (!h) return Fail("null handle");
  if (family < 0 || family >= kNumFamilies) return Fail("bad family");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  DrainProfile(h);
  if (ms) *ms = h->famMs[family];
  if (launches) *launches = h->famLaunches[family];
  return 0;
}

}  // extern "C"
#endif  // !AITHER_EQ_TU
