// kernels.cuh -- the sm_100a kernels of the hot path.
//
// All kernels are fp64, bound by HBM bandwidth or the FP64 pipe (no tensor-core work: nothing here
// is GEMM shaped). One thread owns one cell; face fluxes are computed once per face and shared
// between the two owning cells through shared memory (owner-writes, no atomics).
//
// Kernel <-> reference map:
//   BcKernel        K11  procBlock::AssignInviscidGhostCells       src/procBlock.cpp:2449
//   ResidualKernel  K1   procBlock::CalcInvFluxI/J/K + reset       src/procBlock.cpp:384,522,660,953
//   PrepKernel      K5   CalcBlockTimeStep + AddDiagonalTerms + Invert + InitializeMatrixUpdate
//                        src/procBlock.cpp:798, src/linearSolver.cpp:146,177,111
//   DplurKernel     K6   dplur::DPLUR                               src/linearSolver.cpp:473
//   LusgsKernel     K7   lusgs::LUSGS_Forward/Backward              src/linearSolver.cpp:341,385
//   AxmbKernel      K8   linearSolver::AXmB/Residual + L2           src/linearSolver.cpp:58,92
//   UpdateKernel    K9   procBlock::UpdateBlock                     src/procBlock.cpp:826
//   StoreOldKernel  K10  procBlock::AssignSolToTimeN                src/procBlock.cpp:1037
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"
#include "physics.cuh"
#include "turbulence.cuh"

namespace aither {

// scalar solver parameters, passed by value
struct Params {
  Gas gas;
  Transport tr;
  double kappa, theta, zeta, relax, dualTimeCFL, dtNondim, viscCFLCoeff;
  int isMultilevelTime;
  int matrixRequiresInit;
  int wenoZ;
  int isViscous;
  int viscRecon;   // 0 central, 1 centralFourth
};

// per-iteration reduction results (device + pinned host mirror)
struct IterResult {
  double l2[AITHER_MAX_SPECIES + 6];
  double matrixSumSq;
  double linf;
  int linfBlock, linfI, linfJ, linfK, linfEqn;
  int pad;
};

template <int NEQ>
__device__ __forceinline__ void LoadCell(const double *__restrict__ f, long long fs, long long idx,
                                         double *s) {
#pragma unroll
  for (int e = 0; e < NEQ; ++e) s[e] = __ldg(f + e * fs + idx);
}
template <int NEQ>
__device__ __forceinline__ void StoreCell(double *__restrict__ f, long long fs, long long idx,
                                          const double *s) {
#pragma unroll
  for (int e = 0; e < NEQ; ++e) f[e * fs + idx] = s[e];
}

// ---------------------------------------------------------------------------------------------
// layout conversion: reference array-of-structs (host order) <-> device structure-of-arrays
// src extents (SI,SJ,SK) with `nc` doubles per entry; entry (ii,jj,kk) maps to device index
// (ii+oi) + (jj+oj)*sj + (kk+ok)*sk
__global__ void AosToSoaKernel(const double *__restrict__ src, int SI, int SJ, int SK, int nc,
                               double *__restrict__ dst, long long fs, int oi, int oj, int ok,
                               int sj, long long sk) {
  const long long n = static_cast<long long>(SI) * SJ * SK;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ii = static_cast<int>(t % SI);
    const int jj = static_cast<int>((t / SI) % SJ);
    const int kk = static_cast<int>(t / (static_cast<long long>(SI) * SJ));
    const long long d = (ii + oi) + static_cast<long long>(jj + oj) * sj + (kk + ok) * sk;
    for (int c = 0; c < nc; ++c) dst[c * fs + d] = src[t * nc + c];
  }
}
__global__ void SoaToAosKernel(double *__restrict__ dstAos, int SI, int SJ, int SK, int nc,
                               const double *__restrict__ src, long long fs, int oi, int oj,
                               int ok, int sj, long long sk) {
  const long long n = static_cast<long long>(SI) * SJ * SK;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ii = static_cast<int>(t % SI);
    const int jj = static_cast<int>((t / SI) % SJ);
    const int kk = static_cast<int>(t / (static_cast<long long>(SI) * SJ));
    const long long d = (ii + oi) + static_cast<long long>(jj + oj) * sj + (kk + ok) * sk;
    for (int c = 0; c < nc; ++c) dstAos[t * nc + c] = src[c * fs + d];
  }
}

// ---------------------------------------------------------------------------------------------
// K11 boundary-condition ghost fill. One thread per (boundary face, ghost layer).
struct SurfDev {
  int type, surfType, tag, bcIndex;
  int lo[3], hi[3];      // cell ranges; the normal direction has lo = boundary face index
  long long faceOffset;  // prefix sum of faces * layers over the surfaces of the block
};

template <int NS, int NT>
__global__ void BcKernel(BlockDev b, Params p, const SurfDev *__restrict__ surfs, int nsurf,
                         const aither_bc_state *__restrict__ bcs, long long total) {
  using E = Eq<NS, NT>;
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t >= total) return;
  int s = 0;
  while (s + 1 < nsurf && surfs[s + 1].faceOffset <= t) ++s;
  const SurfDev sf = surfs[s];
  const int d3 = (sf.surfType - 1) / 2;
  const int d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
  const int n1 = sf.hi[d1] - sf.lo[d1], n2 = sf.hi[d2] - sf.lo[d2];
  long long r = t - sf.faceOffset;
  const int a1 = static_cast<int>(r % n1);
  r /= n1;
  const int a2 = static_cast<int>(r % n2);
  const int layer = static_cast<int>(r / n2) + 1;
  const int nd[3] = {b.ni, b.nj, b.nk};
  const int r3 = sf.lo[d3];
  // ref: src/procBlock.cpp:2470-2486
  int gCell, iCell, aCell;
  if (sf.surfType % 2 == 0) {
    gCell = r3 + layer - 1;
    iCell = max(r3 - layer, 0);
    aCell = r3 - 1;
  } else {
    gCell = r3 - layer;
    iCell = min(r3 + layer - 1, nd[d3] - 1);
    aCell = r3;
  }
  int bcType = sf.type;
  if (bcType == AITHER_BC_VISCOUS_WALL) bcType = AITHER_BC_SLIP_WALL;
  int c[3];
  c[d1] = sf.lo[d1] + a1;
  c[d2] = sf.lo[d2] + a2;
  c[d3] = bcType == AITHER_BC_SLIP_WALL ? iCell : aCell;
  double interior[E::neq], ghost[E::neq], area[3];
  LoadCell<E::neq>(b.state, b.fs, CellIdx(b, c[0], c[1], c[2]), interior);
  c[d3] = r3;
  const long long fidx = CellIdx(b, c[0], c[1], c[2]);
#pragma unroll
  for (int q = 0; q < 3; ++q) area[q] = __ldg(b.fA[d3] + q * b.fs + fidx);
  GhostState<NS, NT>(p.gas, interior, bcType, area, sf.surfType, bcs[sf.bcIndex], layer, ghost,
                     &p.tr);
  c[d3] = gCell;
  StoreCell<E::neq>(b.state, b.fs, CellIdx(b, c[0], c[1], c[2]), ghost);
}

// ---------------------------------------------------------------------------------------------
// K1 residual assembly: reconstruction + Riemann flux fused, I/J/K sweeps in one kernel.
constexpr int kTI = 32, kTJ = 4, kTK = 2;
constexpr int kResThreads = kTI * kTJ * kTK;
constexpr int kFaceMax = (kTI + 1) * kTJ * kTK > kTI * (kTJ + 1) * kTK
                             ? ((kTI + 1) * kTJ * kTK > kTI * kTJ * (kTK + 1) ? (kTI + 1) * kTJ * kTK
                                                                              : kTI * kTJ * (kTK + 1))
                             : (kTI * (kTJ + 1) * kTK > kTI * kTJ * (kTK + 1) ? kTI * (kTJ + 1) * kTK
                                                                              : kTI * kTJ * (kTK + 1));

template <int D>
__device__ __forceinline__ int LocalFace(int lx, int ly, int lz) {
  if (D == 0) return lx + (kTI + 1) * (ly + kTJ * lz);
  if (D == 1) return lx + kTI * (ly + (kTJ + 1) * lz);
  return lx + kTI * (ly + kTJ * lz);
}

// flux through one face times its area; `idx` = index of the cell on the upper side of the face
template <int NS, int NT, int RECON, int LIM, int FLUX>
__device__ __forceinline__ void FaceFlux(const BlockDev &b, const Params &p, int d, long long idx,
                                         double *out) {
  using E = Eq<NS, NT>;
  const long long st = Stride(b, d);
  double fl[E::neq], fr[E::neq];
  if (RECON == AITHER_RECON_CONSTANT) {
    LoadCell<E::neq>(b.state, b.fs, idx - st, fl);
    LoadCell<E::neq>(b.state, b.fs, idx, fr);
  } else if (RECON == AITHER_RECON_MUSCL) {
    // ref: src/procBlock.cpp:406-418
    double um2[E::neq], um1[E::neq], u0[E::neq], up1[E::neq];
    LoadCell<E::neq>(b.state, b.fs, idx - 2 * st, um2);
    LoadCell<E::neq>(b.state, b.fs, idx - st, um1);
    LoadCell<E::neq>(b.state, b.fs, idx, u0);
    LoadCell<E::neq>(b.state, b.fs, idx + st, up1);
    const double *cw = b.cw[d];
    const double wm2 = __ldg(cw + idx - 2 * st), wm1 = __ldg(cw + idx - st), w0 = __ldg(cw + idx),
                 wp1 = __ldg(cw + idx + st);
    Muscl<E::neq, LIM>(um2, um1, u0, p.kappa, wm2, wm1, w0, fl);
    Muscl<E::neq, LIM>(up1, u0, um1, p.kappa, wp1, w0, wm1, fr);
  } else {
    // ref: src/procBlock.cpp:420-436
    double u[6][E::neq], w[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      LoadCell<E::neq>(b.state, b.fs, idx + (o - 3) * st, u[o]);
      w[o] = __ldg(b.cw[d] + idx + (o - 3) * st);
    }
    {
      const double wl[5] = {w[0], w[1], w[2], w[3], w[4]};
      const WenoGeom g = WenoSetup(wl);
#pragma unroll
      for (int e = 0; e < E::neq; ++e)
        fl[e] = p.wenoZ ? Weno1<true>(g, u[0][e], u[1][e], u[2][e], u[3][e], u[4][e])
                        : Weno1<false>(g, u[0][e], u[1][e], u[2][e], u[3][e], u[4][e]);
    }
    {
      const double wr[5] = {w[5], w[4], w[3], w[2], w[1]};
      const WenoGeom g = WenoSetup(wr);
#pragma unroll
      for (int e = 0; e < E::neq; ++e)
        fr[e] = p.wenoZ ? Weno1<true>(g, u[5][e], u[4][e], u[3][e], u[2][e], u[1][e])
                        : Weno1<false>(g, u[5][e], u[4][e], u[3][e], u[2][e], u[1][e]);
    }
  }
  double area[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) area[q] = __ldg(b.fA[d] + q * b.fs + idx);
  double flux[E::neq];
  InviscidFlux<NS, NT, FLUX>(p.gas, fl, fr, area, flux);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) out[e] = flux[e] * area[3];
}

template <int NS, int NT, int RECON, int LIM, int FLUX, int D>
__device__ __forceinline__ void ResidualPass(const BlockDev &b, const Params &p,
                                             double (*sflux)[kFaceMax], int i0, int j0, int k0,
                                             int tx, int ty, int tz, int tid, bool cellValid,
                                             long long idx, const double *s, double sos,
                                             double *res, double &specRad) {
  using E = Eq<NS, NT>;
  const int nd[3] = {b.ni, b.nj, b.nk};
  {
    // the face on the lower side of this thread's cell (also for the cell one past the end,
    // whose lower face is the block's upper boundary face)
    const int g[3] = {i0 + tx, j0 + ty, k0 + tz};
    bool valid = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) valid = valid && (q == D ? g[q] <= nd[q] : g[q] < nd[q]);
    if (valid) {
      double f[E::neq];
      FaceFlux<NS, NT, RECON, LIM, FLUX>(b, p, D, idx, f);
      const int lf = LocalFace<D>(tx, ty, tz);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) sflux[e][lf] = f[e];
    }
  }
  constexpr int ext = D == 0 ? kTJ * kTK : (D == 1 ? kTI * kTK : kTI * kTJ);
  if (tid < ext) {
    // faces on the upper boundary of the tile
    int l[3];
    if (D == 0) {
      l[0] = kTI; l[1] = tid % kTJ; l[2] = tid / kTJ;
    } else if (D == 1) {
      l[0] = tid % kTI; l[1] = kTJ; l[2] = tid / kTI;
    } else {
      l[0] = tid % kTI; l[1] = tid / kTI; l[2] = kTK;
    }
    const int g[3] = {i0 + l[0], j0 + l[1], k0 + l[2]};
    bool valid = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) valid = valid && (q == D ? g[q] <= nd[q] : g[q] < nd[q]);
    if (valid) {
      double f[E::neq];
      FaceFlux<NS, NT, RECON, LIM, FLUX>(b, p, D, CellIdx(b, g[0], g[1], g[2]), f);
      const int lf = LocalFace<D>(l[0], l[1], l[2]);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) sflux[e][lf] = f[e];
    }
  }
  __syncthreads();
  if (cellValid) {
    // ref accumulation order (src/procBlock.cpp:447-463): the lower face subtracts first, then
    // the upper face adds
    const int lo = LocalFace<D>(tx, ty, tz);
    const int hi = LocalFace<D>(tx + (D == 0), ty + (D == 1), tz + (D == 2));
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      res[e] -= sflux[e][lo];
      res[e] += sflux[e][hi];
    }
    double fL[4], fR[4];
    const long long st = Stride(b, D);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      fL[q] = __ldg(b.fA[D] + q * b.fs + idx);
      fR[q] = __ldg(b.fA[D] + q * b.fs + idx + st);
    }
    specRad += InvCellSpectralRadius<NS>(s, sos, fL, fR);  // ref: :468-488
  }
  __syncthreads();
}

template <int NS, int NT, int RECON, int LIM, int FLUX>
__global__ void __launch_bounds__(kResThreads)
    ResidualKernel(BlockDev b, Params p, int implicitScalar) {
  using E = Eq<NS, NT>;
  __shared__ double sflux[E::neq][kFaceMax];
  const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
  const int tid = tx + kTI * (ty + kTJ * tz);
  const int i0 = blockIdx.x * kTI, j0 = blockIdx.y * kTJ, k0 = blockIdx.z * kTK;
  const int i = i0 + tx, j = j0 + ty, k = k0 + tz;
  const bool cellValid = i < b.ni && j < b.nj && k < b.nk;
  const long long idx = CellIdx(b, i, j, k);
  double res[E::neq], s[E::neq];
  double specRad = 0.0, sos = 0.0;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) res[e] = 0.0;
  if (cellValid) {
    LoadCell<E::neq>(b.state, b.fs, idx, s);
    sos = SoS<NS>(p.gas, s);
  }
  ResidualPass<NS, NT, RECON, LIM, FLUX, 0>(b, p, sflux, i0, j0, k0, tx, ty, tz, tid, cellValid,
                                            idx, s, sos, res, specRad);
  ResidualPass<NS, NT, RECON, LIM, FLUX, 1>(b, p, sflux, i0, j0, k0, tx, ty, tz, tid, cellValid,
                                            idx, s, sos, res, specRad);
  ResidualPass<NS, NT, RECON, LIM, FLUX, 2>(b, p, sflux, i0, j0, k0, tx, ty, tz, tid, cellValid,
                                            idx, s, sos, res, specRad);
  if (cellValid) {
    StoreCell<E::neq>(b.resid, b.fs, idx, res);
    b.specRad[idx] = specRad;
    b.specRad[b.fs + idx] = 0.0;
    // scalar implicit diagonal accumulates the same spectral radii (ref: :485-488); the
    // diagonal was zeroed by ResetDiagonal, so the sum starts from 0 exactly as specRadius_
    if (implicitScalar) b.diag[idx] = specRad;
  }
}

// ---------------------------------------------------------------------------------------------
// K5 time step, diagonal, inverse, right-hand side b and initial update x0.
enum PrepBits { kPrepDt = 1, kPrepDiag = 2, kPrepInit = 4 };

template <int NS, int NT, bool LOCAL_RES = false>
__device__ __forceinline__ void RhsB(const BlockDev &b, const Params &p, long long idx,
                                     const double *s, double vol, double dt, double *out,
                                     const double *resLocal = nullptr) {
  // b = -R/theta + SolDeltaNm1 - SolDeltaMmN; ref: src/procBlock.cpp:1010-1034,
  // src/linearSolver.cpp:124-129
  using E = Eq<NS, NT>;
  double cons[E::neq];
  PrimToCons<NS, NT>(p.gas, s, cons);
  const double thetaInv = 1.0 / p.theta;
  const double coeff = (vol * (1.0 + p.zeta)) / (dt * p.theta);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) {
    const double cn = __ldg(b.consN + e * b.fs + idx);
    double nm1 = 0.0;
    if (p.isMultilevelTime) {
      const double c1 = (vol * p.zeta) / (dt * p.theta);
      nm1 = c1 * (cn - __ldg(b.consNm1 + e * b.fs + idx));
    }
    const double mmn = coeff * (cons[e] - cn);
    const double r = LOCAL_RES ? resLocal[e] : __ldg(b.resid + e * b.fs + idx);
    out[e] = -thetaInv * r + nm1 - mmn;
  }
}

template <int NS, int NT>
__global__ void __launch_bounds__(256) PrepKernel(BlockDev b, Params p, double cfl, int bits) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  const double vol = __ldg(b.vol + idx);
  const double srF = b.specRad[idx], srT = b.specRad[b.fs + idx];
  const double srMax = fmax(srF, srT);
  double dt;
  if (bits & kPrepDt) {
    // ref: src/procBlock.cpp:782-821
    dt = p.dtNondim > 0.0 ? p.dtNondim : cfl * (vol / srMax);
    b.dt[idx] = dt;
  } else {
    dt = b.dt[idx];
  }
  double dinv, dinvT = 0.0;
  if (bits & kPrepDiag) {
    // ref: src/linearSolver.cpp:146-188
    double diagVolTime = (vol * (1.0 + p.zeta)) / (dt * p.theta);
    if (p.dualTimeCFL > 0.0) diagVolTime += srMax / p.dualTimeCFL;
    double a = b.diag[idx];
    a *= p.relax;
    a += diagVolTime;
    b.diag[idx] = a;
    dinv = 1.0 / a;
    b.dinv[idx] = dinv;
    if (NT > 0) {  // uncoupled scalar diagonal {flow, turbulence}
      double at = b.diag[b.fs + idx];
      at *= p.relax;
      at += diagVolTime;
      b.diag[b.fs + idx] = at;
      dinvT = 1.0 / at;
      b.dinv[b.fs + idx] = dinvT;
    }
  } else {
    dinv = b.dinv[idx];
    if (NT > 0) dinvT = b.dinv[b.fs + idx];
  }
  if (bits & kPrepInit) {
    double s[E::neq], rb[E::neq];
    LoadCell<E::neq>(b.state, b.fs, idx, s);
    RhsB<NS, NT>(b, p, idx, s, vol, dt, rb);
    StoreCell<E::neq>(b.rhs, b.fs, idx, rb);
    // ref: src/linearSolver.cpp:111-144 (x = D^-1 b when the solver needs initialisation, else 0)
#pragma unroll
    for (int e = 0; e < E::neq; ++e)
      b.x[e * b.fs + idx] = p.matrixRequiresInit ? rb[e] * (e < NS + 4 ? dinv : dinvT) : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// implicit off-diagonal sums L and U for one cell; ref: src/procBlock.cpp:1056-1170
__device__ __forceinline__ bool ConnAcross(const BlockDev &b, int surf, int c1, int n1, int c2) {
  const uint8_t *m = b.connFace[surf - 1];
  return m != nullptr && m[c1 + n1 * c2] != 0;
}

// viscous parts of the face spectral radii of neighbour cell `nidx` (state sn): flow
// (ViscFaceSpectralRadius, include/spectralRadius.hpp:126-151) and turbulence equations
// (turbModel::ViscousFaceSpectralRadius, src/turbulence.cpp:513-527,796-808), from the neighbour's
// stored viscosity, eddy viscosity and blending function (ref: src/procBlock.cpp:1069-1076)
template <int NS, int NT>
__device__ __forceinline__ void NeighbourViscTerms(const BlockDev &b, const Params &p,
                                                   const double *sn, long long nidx, double length,
                                                   double *extra, double *extraT) {
  const double rho = SpeciesSum<NS>(sn);
  const double mu = __ldg(b.viscosity + nidx);
  const double mut = NT > 0 ? __ldg(b.eddyVisc + nidx) : 0.0;
  *extra = length * ViscSpecFactor(p.tr, rho, Gamma<NS>(p.gas, sn), mu, mut);
  if (NT > 0)
    *extraT = length * TurbViscSpecFactor(p.tr.turbModel, p.tr.scaling, rho, sn[NS + 4],
                                          sn[NS + 4 + (NT > 1 ? 1 : 0)], mu, mut,
                                          __ldg(b.f1 + nidx));
}

template <int NS, int NT, bool LOWER, bool UPPER>
__device__ __forceinline__ void OffDiagonals(const BlockDev &b, const Params &p,
                                             const double *__restrict__ x, int i, int j, int k,
                                             long long idx, double *L, double *U) {
  using E = Eq<NS, NT>;
  const int c[3] = {i, j, k};
  const int nd[3] = {b.ni, b.nj, b.nk};
#pragma unroll
  for (int e = 0; e < E::neq; ++e) {
    L[e] = 0.0;
    U[e] = 0.0;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const long long st = Stride(b, d);
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    if (LOWER) {
      if (c[d] > 0 || ConnAcross(b, 2 * d + 1, c[d1], nd[d1], c[d2])) {
        double sn[E::neq], dun[E::neq], fa[4], od[E::neq];
        LoadCell<E::neq>(b.state, b.fs, idx - st, sn);
        LoadCell<E::neq>(x, b.fs, idx - st, dun);
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[d] + q * b.fs + idx);
        double extra = 0.0, extraT = 0.0;
        if (p.isViscous)
          NeighbourViscTerms<NS, NT>(b, p, sn, idx - st, fa[3] / __ldg(b.dist[d] + idx), &extra,
                                     &extraT);
        OffDiagScalar<NS, NT>(p.gas, sn, dun, fa, true, od, extra, extraT);
#pragma unroll
        for (int e = 0; e < E::neq; ++e) L[e] += od[e];
      }
    }
    if (UPPER) {
      if (c[d] < nd[d] - 1 || ConnAcross(b, 2 * d + 2, c[d1], nd[d1], c[d2])) {
        double sn[E::neq], dun[E::neq], fa[4], od[E::neq];
        LoadCell<E::neq>(b.state, b.fs, idx + st, sn);
        LoadCell<E::neq>(x, b.fs, idx + st, dun);
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[d] + q * b.fs + idx + st);
        double extra = 0.0, extraT = 0.0;
        if (p.isViscous)
          NeighbourViscTerms<NS, NT>(b, p, sn, idx + st, fa[3] / __ldg(b.dist[d] + idx + st),
                                     &extra, &extraT);
        OffDiagScalar<NS, NT>(p.gas, sn, dun, fa, false, od, extra, extraT);
#pragma unroll
        for (int e = 0; e < E::neq; ++e) U[e] += od[e];
      }
    }
  }
}

// K6 DPLUR (Jacobi) sweep: xout = D^-1 (b + L(xin) - U(xin)); ref: src/linearSolver.cpp:473-507
template <int NS, int NT>
__global__ void __launch_bounds__(256)
    DplurKernel(BlockDev b, Params p, const double *__restrict__ xin, double *__restrict__ xout) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double L[E::neq], U[E::neq];
  OffDiagonals<NS, NT, true, true>(b, p, xin, i, j, k, idx, L, U);
  const double dinvF = __ldg(b.dinv + idx);
  const double dinvT = NT > 0 ? __ldg(b.dinv + b.fs + idx) : 0.0;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) {
    const double rb = __ldg(b.rhs + e * b.fs + idx);
    xout[e * b.fs + idx] = ((rb + 0.0) + (L[e] - U[e])) * (e < NS + 4 ? dinvF : dinvT);
  }
}

// K7 LU-SGS along one i+j+k hyperplane; ref: src/linearSolver.cpp:341-428. Cells of a plane do
// not couple, so any order inside the plane reproduces the reference's lexicographic result.
template <int NS, int NT, bool FORWARD>
__global__ void __launch_bounds__(128)
    LusgsPlaneKernel(BlockDev b, Params p, int plane, int fullGS) {
  using E = Eq<NS, NT>;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= b.nj || k >= b.nk) return;
  const int i = plane - j - k;
  if (i < 0 || i >= b.ni) return;
  const long long idx = CellIdx(b, i, j, k);
  const double dinvF = __ldg(b.dinv + idx);
  const double dinvT = NT > 0 ? __ldg(b.dinv + b.fs + idx) : 0.0;
#define dinv (e < NS + 4 ? dinvF : dinvT)
  double L[E::neq], U[E::neq];
  if (FORWARD) {
    if (fullGS) {
      OffDiagonals<NS, NT, true, true>(b, p, b.x, i, j, k, idx, L, U);
    } else {
      OffDiagonals<NS, NT, true, false>(b, p, b.x, i, j, k, idx, L, U);
    }
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double rb = __ldg(b.rhs + e * b.fs + idx);
      b.x[e * b.fs + idx] = (rb + (L[e] - U[e])) * dinv;
    }
  } else {
    if (fullGS) {
      OffDiagonals<NS, NT, true, true>(b, p, b.x, i, j, k, idx, L, U);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) {
        const double rb = __ldg(b.rhs + e * b.fs + idx);
        b.x[e * b.fs + idx] = ((rb + L[e]) - U[e]) * dinv;
      }
    } else {
      OffDiagonals<NS, NT, false, true>(b, p, b.x, i, j, k, idx, L, U);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) {
        const double xo = b.x[e * b.fs + idx];
        b.x[e * b.fs + idx] = xo - U[e] * dinv;
      }
    }
  }
#undef dinv
}

// ---------------------------------------------------------------------------------------------
// deterministic block reductions: warp shuffle tree, then one partial per block
__device__ __forceinline__ double WarpSum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void BlockSumToPartials(double *vals, double *partials, int blockLinear,
                                                   int tid, int nthreads) {
  __shared__ double sh[NV][32];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double w = WarpSum(vals[v]);
    if (lane == 0) sh[v][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (nthreads + 31) >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double w = lane < nw ? sh[v][lane] : 0.0;
      w = WarpSum(w);
      if (lane == 0) partials[static_cast<long long>(blockLinear) * NV + v] = w;
    }
  }
}

// K8 matrix residual f - (A x - (L - U) - b) and its sum of squares;
// ref: src/linearSolver.cpp:58-109, src/mgSolution.cpp:198-206
template <int NS, int NT>
__global__ void __launch_bounds__(256)
    AxmbKernel(BlockDev b, Params p, double *__restrict__ partials, int storeField) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  double sq = 0.0;
  if (i < b.ni && j < b.nj) {
    const long long idx = CellIdx(b, i, j, k);
    double L[E::neq], U[E::neq];
    OffDiagonals<NS, NT, true, true>(b, p, b.x, i, j, k, idx, L, U);
    const double aF = __ldg(b.diag + idx);
    const double aT = NT > 0 ? __ldg(b.diag + b.fs + idx) : 0.0;
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double rb = __ldg(b.rhs + e * b.fs + idx);
      const double ax = b.x[e * b.fs + idx] * (e < NS + 4 ? aF : aT);
      const double mr = 0.0 - ((ax - (L[e] - U[e])) - rb);
      if (storeField) b.mres[e * b.fs + idx] = mr;
      sq += mr * mr;
    }
  }
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  const int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  BlockSumToPartials<1>(&sq, partials, blockLinear, tid, blockDim.x * blockDim.y);
}

// K9 state update + residual norms; ref: src/procBlock.cpp:826-871, :902-915
struct LinfCand {
  double v;
  long long key;  // traversal order ((k*nj + j)*ni + i)*neq + e; smaller wins ties
};
__device__ __forceinline__ LinfCand LinfBetter(LinfCand a, LinfCand c) {
  return (c.v > a.v || (c.v == a.v && c.key < a.key)) ? c : a;
}

template <int NS, int NT>
__global__ void __launch_bounds__(256)
    UpdateKernel(BlockDev b, Params p, double *__restrict__ partials,
                 LinfCand *__restrict__ linfPartials) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  double sq[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) sq[e] = 0.0;
  LinfCand best;
  best.v = 0.0;  // the reference starts from linf = 0 and uses a strict '>' (resid.hpp:33)
  best.key = 0x7fffffffffffffffLL;
  if (i < b.ni && j < b.nj) {
    const long long idx = CellIdx(b, i, j, k);
    double s[E::neq], du[E::neq], sn[E::neq];
    LoadCell<E::neq>(b.state, b.fs, idx, s);
    LoadCell<E::neq>(b.x, b.fs, idx, du);
    UpdatePrimWithCons<NS, NT>(p.gas, s, du, sn);
    StoreCell<E::neq>(b.state, b.fs, idx, sn);
    const long long cellKey =
        ((static_cast<long long>(k) * b.nj + j) * b.ni + i) * static_cast<long long>(E::neq);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double r = __ldg(b.resid + e * b.fs + idx);
      sq[e] = r * r;
      LinfCand c;
      c.v = r;
      c.key = cellKey + e;
      if (r > best.v) best = c;  // increasing e: first maximum wins, like the reference loop
    }
  }
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  const int nthreads = blockDim.x * blockDim.y;
  const int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  BlockSumToPartials<E::neq>(sq, partials, blockLinear, tid, nthreads);
  // arg-max
  __shared__ LinfCand shc[32];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    LinfCand c;
    c.v = __shfl_down_sync(0xffffffffu, best.v, o);
    c.key = __shfl_down_sync(0xffffffffu, best.key, o);
    best = LinfBetter(best, c);
  }
  if (lane == 0) shc[warp] = best;
  __syncthreads();
  if (tid == 0) {
    LinfCand r = shc[0];
    for (int w = 1; w < (nthreads + 31) / 32; ++w) r = LinfBetter(r, shc[w]);
    linfPartials[blockLinear] = r;
  }
}

// final pass over the per-block partials of one procBlock, accumulating into the iteration's
// result record in a fixed order (deterministic run to run)
constexpr int kFinalThreads = 512;
__global__ void __launch_bounds__(kFinalThreads)
    FinalizeSumKernel(const double *__restrict__ partials, int nPartials, int nv,
                      double *__restrict__ out) {
  // one block per value; each thread strides the partial list, then a fixed shuffle/shared tree
  __shared__ double sh[kFinalThreads / 32];
  const int v = blockIdx.x;
  double acc = 0.0;
  for (int q = threadIdx.x; q < nPartials; q += kFinalThreads)
    acc += partials[static_cast<long long>(q) * nv + v];
  acc = WarpSum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double w = threadIdx.x < kFinalThreads / 32 ? sh[threadIdx.x] : 0.0;
    w = WarpSum(w);
    if (threadIdx.x == 0) out[v] += w;
  }
}
__global__ void __launch_bounds__(kFinalThreads)
    FinalizeLinfKernel(const LinfCand *__restrict__ cands, int n, BlockDev b, int neq,
                       IterResult *__restrict__ res) {
  __shared__ LinfCand sh[kFinalThreads / 32];
  LinfCand best;
  best.v = 0.0;
  best.key = 0x7fffffffffffffffLL;
  for (int q = threadIdx.x; q < n; q += kFinalThreads) best = LinfBetter(best, cands[q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    LinfCand c;
    c.v = __shfl_down_sync(0xffffffffu, best.v, o);
    c.key = __shfl_down_sync(0xffffffffu, best.key, o);
    best = LinfBetter(best, c);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kFinalThreads / 32; ++w) best = LinfBetter(best, sh[w]);
  }
  if (threadIdx.x == 0 && best.v > res->linf) {  // strict: earlier blocks win ties
    long long key = best.key;
    const int e = static_cast<int>(key % neq);
    key /= neq;
    res->linf = best.v;
    res->linfBlock = b.parentBlock;
    res->linfI = static_cast<int>(key % b.ni);
    key /= b.ni;
    res->linfJ = static_cast<int>(key % b.nj);
    res->linfK = static_cast<int>(key / b.nj);
    res->linfEqn = e + 1;
  }
}

// K10 U^n <- cons(state); ref: src/procBlock.cpp:1037-1053
template <int NS, int NT>
__global__ void __launch_bounds__(256) StoreOldKernel(BlockDev b, Params p, int copyToNm1) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double s[E::neq], c[E::neq];
  LoadCell<E::neq>(b.state, b.fs, idx, s);
  PrimToCons<NS, NT>(p.gas, s, c);
  StoreCell<E::neq>(b.consN, b.fs, idx, c);
  if (copyToNm1) StoreCell<E::neq>(b.consNm1, b.fs, idx, c);
}

__global__ void FillKernel(double *__restrict__ p, long long n, double v) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x)
    p[t] = v;
}

}  // namespace aither
