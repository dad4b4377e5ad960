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
#include "blockjac.cuh"

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
  // one nonlinear iteration per step of a single-level scheme: U^m = U^n at the only iteration,
  // so the time terms of b vanish identically and U^n is never read (nor stored)
  int timeTermsVanish;
  // plane-marching kernels: L2 prefetch of the next plane's register-fed operands
  int prefetch;
};

__device__ __forceinline__ void PrefetchL2(const void *ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

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
static __global__ void AosToSoaKernel(const double *__restrict__ src, int SI, int SJ, int SK, int nc,
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
static __global__ void SoaToAosKernel(double *__restrict__ dstAos, int SI, int SJ, int SK, int nc,
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

// ghost shell of `src` into `dst` (nc fields). With the state update fused into the matrix-residual
// pass the two state buffers change roles every iteration; the ghost cells the reference's state_
// holds after an iteration (filled at the START of that iteration, src/gridLevel.cpp:287-319) then
// sit in the buffer that has just become the old one. Nothing on the path reads ghost cells before
// the next fill, so they are only moved over when the state is downloaded.
static __global__ void GhostShellCopyKernel(BlockDev b, const double *__restrict__ src,
                                            double *__restrict__ dst, int nc) {
  const int NI = b.ni + 2 * b.g, NJ = b.nj + 2 * b.g, NK = b.nk + 2 * b.g;
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t >= static_cast<long long>(NI) * NJ * NK) return;
  const int i = static_cast<int>(t % NI) - b.g, j = static_cast<int>((t / NI) % NJ) - b.g;
  const int k = static_cast<int>(t / (static_cast<long long>(NI) * NJ)) - b.g;
  if (i >= 0 && i < b.ni && j >= 0 && j < b.nj && k >= 0 && k < b.nk) return;
  const long long idx = CellIdx(b, i, j, k);
  for (int e = 0; e < nc; ++e) dst[e * b.fs + idx] = src[e * b.fs + idx];
}

// ---------------------------------------------------------------------------------------------
// K11 boundary-condition ghost fill. One thread per (boundary face, ghost layer).
struct SurfDev {
  int type, surfType, tag, bcIndex;
  int lo[3], hi[3];      // cell ranges; the normal direction has lo = boundary face index
  long long faceOffset;  // prefix sum of faces * layers over the surfaces of the block
};

// average and maximum outward Mach number of the boundary-adjacent cells of every non-reflecting
// inlet / outlet patch (ref: src/procBlock.cpp:6235-6261): one thread block per surface,
// fixed-order tree reduction (run-to-run identical); out[2 s] = average, out[2 s + 1] = maximum
template <int NS, int NT>
__global__ void __launch_bounds__(256) PatchMachKernel(BlockDev b, Params p,
                                                       const SurfDev *__restrict__ surfs,
                                                       const aither_bc_state *__restrict__ bcs,
                                                       double *__restrict__ out) {
  using E = Eq<NS, NT>;
  const SurfDev sf = surfs[blockIdx.x];
  if (!(sf.type == AITHER_BC_INLET || sf.type == AITHER_BC_PRESSURE_OUTLET) ||
      !bcs[sf.bcIndex].isNonreflecting)
    return;
  const int d3 = (sf.surfType - 1) / 2;
  const int d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
  const int n1 = sf.hi[d1] - sf.lo[d1], n2 = sf.hi[d2] - sf.lo[d2];
  const int r3 = sf.lo[d3];
  const bool isLower = sf.surfType % 2 == 1;
  const int aCell = isLower ? r3 : r3 - 1;
  double sum = 0.0, mx = -1.7976931348623157e308;
  for (int t = threadIdx.x; t < n1 * n2; t += blockDim.x) {
    int c[3];
    c[d1] = sf.lo[d1] + t % n1;
    c[d2] = sf.lo[d2] + t / n1;
    c[d3] = aCell;
    double s[E::neq];
    LoadCell<E::neq>(b.state, b.fs, CellIdx(b, c[0], c[1], c[2]), s);
    c[d3] = r3;
    const long long fidx = CellIdx(b, c[0], c[1], c[2]);
    double vn = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double a = __ldg(b.fA[d3] + q * b.fs + fidx);
      vn += s[E::imx + q] * (isLower ? -1.0 * a : a);
    }
    const double mach = vn / SoS<NS>(p.gas, s);
    sum += mach;
    mx = fmax(mx, mach);
  }
  __shared__ double ssum[256], smax[256];
  ssum[threadIdx.x] = sum;
  smax[threadIdx.x] = mx;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      ssum[threadIdx.x] += ssum[threadIdx.x + w];
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + w]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = ssum[0] / static_cast<double>(n1 * n2);
    out[2 * blockIdx.x + 1] = smax[0];
  }
}

template <int NS, int NT>
__global__ void BcKernel(BlockDev b, Params p, const SurfDev *__restrict__ surfs, int nsurf,
                         const aither_bc_state *__restrict__ bcs, long long total,
                         const double *__restrict__ patchMach = nullptr) {
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
  // consecutive threads along i wherever the surface has an i extent (j-surfaces: direction 2 is
  // i, k-surfaces: direction 1): their loads and stores then fall on consecutive addresses
  int a1, a2;
  if (d3 == 1) {
    a2 = static_cast<int>(r % n2);
    r /= n2;
    a1 = static_cast<int>(r % n1);
    r /= n1;
  } else {
    a1 = static_cast<int>(r % n1);
    r /= n1;
    a2 = static_cast<int>(r % n2);
    r /= n2;
  }
  const int layer = static_cast<int>(r) + 1;
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
  if (patchMach != nullptr && (bcType == AITHER_BC_INLET || bcType == AITHER_BC_PRESSURE_OUTLET) &&
      bcs[sf.bcIndex].isNonreflecting) {
    // state at time n, time step and gradients of the boundary-adjacent cell as the previous
    // evaluation left them (ref: src/procBlock.cpp:2506-2517)
    BcExtra ex;
    c[d3] = aCell;
    const long long aidx = CellIdx(b, c[0], c[1], c[2]);
    ex.dt = b.dt[aidx];
    double cn[E::neq];
    LoadCell<E::neq>(b.consN, b.fs, aidx, cn);
    ConsToPrim<NS, NT>(p.gas, cn, ex.stateN);
#pragma unroll
    for (int q = 0; q < 3; ++q) ex.pressGrad[q] = b.pressGrad[q * b.fs + aidx];
#pragma unroll
    for (int q = 0; q < 9; ++q) ex.velGrad[q] = b.velGrad[q * b.fs + aidx];
    ex.avgMach = patchMach[2 * s];
    ex.maxMach = patchMach[2 * s + 1];
    GhostState<NS, NT>(p.gas, interior, bcType, area, sf.surfType, bcs[sf.bcIndex], layer, ghost,
                       &p.tr, &ex);
  } else {
    GhostState<NS, NT>(p.gas, interior, bcType, area, sf.surfType, bcs[sf.bcIndex], layer, ghost,
                       &p.tr);
  }
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
                                             double *res, double &specRad, double &specRadT) {
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
    double srT = 0.0;
    specRad += InvCellSpectralRadii<NS>(s, sos, fL, fR, &srT);  // ref: :468-488
    specRadT += srT;
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
  double specRad = 0.0, specRadT = 0.0, sos = 0.0;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) res[e] = 0.0;
  if (cellValid) {
    LoadCell<E::neq>(b.state, b.fs, idx, s);
    sos = SoS<NS>(p.gas, s);
  }
  ResidualPass<NS, NT, RECON, LIM, FLUX, 0>(b, p, sflux, i0, j0, k0, tx, ty, tz, tid, cellValid,
                                            idx, s, sos, res, specRad, specRadT);
  ResidualPass<NS, NT, RECON, LIM, FLUX, 1>(b, p, sflux, i0, j0, k0, tx, ty, tz, tid, cellValid,
                                            idx, s, sos, res, specRad, specRadT);
  ResidualPass<NS, NT, RECON, LIM, FLUX, 2>(b, p, sflux, i0, j0, k0, tx, ty, tz, tid, cellValid,
                                            idx, s, sos, res, specRad, specRadT);
  if (cellValid) {
    StoreCell<E::neq>(b.resid, b.fs, idx, res);
    b.specRad[idx] = specRad;
    b.specRad[b.fs + idx] = NT > 0 ? specRadT : 0.0;
    // scalar implicit diagonal accumulates the same spectral radii (ref: :485-488); the
    // diagonal was zeroed by ResetDiagonal, so the sum starts from 0 exactly as specRadius_
    if (implicitScalar) {
      b.diag[idx] = specRad;
      if (NT > 0) b.diag[b.fs + idx] = specRadT;
    }
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
  const double thetaInv = 1.0 / p.theta;
  if (p.timeTermsVanish) {
    // SolDeltaMmN = coeff (U^m - U^n) = 0 exactly and there is no U^(n-1) term
#pragma unroll
    for (int e = 0; e < E::neq; ++e)
      out[e] = -thetaInv * (LOCAL_RES ? resLocal[e] : __ldg(b.resid + e * b.fs + idx));
    return;
  }
  double cons[E::neq];
  PrimToCons<NS, NT>(p.gas, s, cons);
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

// implicit-matrix flavour: scalar diagonal + Rusanov flux-change off-diagonals (lusgs / dplur),
// full block Jacobians (blusgs / bdplur), or scalar diagonal + Roe flux-change off-diagonals
// (inviscidFluxJacobian: approximateRoe); ref: src/fluxJacobian.cpp:196-238 (OffDiagonal)
enum JacKind { kJacScalar = 0, kJacBlock = 1, kJacRoe = 2 };

// one neighbour's off-diagonal product; `own` = state of the cell being updated (Roe only)
template <int NS, int NT, int JAC>
__device__ __forceinline__ void OffDiagOne(const BlockDev &b, const Params &p, const double *sn,
                                           const double *dun, const double *own, const double *fa,
                                           bool positive, long long nidx, double dist,
                                           double *od) {
  using E = Eq<NS, NT>;
  if constexpr (JAC == kJacBlock) {
    // RusanovBlockOffDiagonal; ref: src/fluxJacobian.cpp:164-194
    double J[Blk<NS, NT>::n];
    RusanovFluxJacobian<NS, NT>(p.gas, sn, fa, positive, J);
    if (p.isViscous) {
      double V[Blk<NS, NT>::n], vg[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) vg[q] = __ldg(b.velGrad + q * b.fs + nidx);
      const double mut = NT > 0 ? __ldg(b.eddyVisc + nidx) : 0.0;
      const double f1 = NT > 0 ? __ldg(b.f1 + nidx) : 0.0;
      ApproxTslJacobian<NS, NT>(p.gas, p.tr, sn, __ldg(b.viscosity + nidx), mut, f1, fa, dist,
                                positive, vg, V);
#pragma unroll
      for (int q = 0; q < Blk<NS, NT>::n; ++q) J[q] = positive ? J[q] - V[q] : J[q] + V[q];
    }
    BlockMult<NS, NT>(J, dun, od);
  } else if constexpr (JAC == kJacRoe) {
    // RoeOffDiagonal; ref: src/fluxJacobian.cpp:240-296. Its caller passes (.., f1, dist, ..) into
    // parameters declared (.., dist, f1, ..), so for viscous flow the reference divides by f1 = 0;
    // only the inviscid use is built (aither_gpu_create refuses approximateRoe + viscous).
    double oldFlux[E::neq], newFlux[E::neq], su[E::neq];
    RoeFlux<NS, NT>(p.gas, sn, own, fa, oldFlux);
    UpdatePrimWithCons<NS, NT>(p.gas, sn, dun, su);
    if (positive) RoeFlux<NS, NT>(p.gas, su, own, fa, newFlux);
    else RoeFlux<NS, NT>(p.gas, own, su, fa, newFlux);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) od[e] = fa[3] * (newFlux[e] - oldFlux[e]);
  } else {
    double extra = 0.0, extraT = 0.0;
    if (p.isViscous) NeighbourViscTerms<NS, NT>(b, p, sn, nidx, fa[3] / dist, &extra, &extraT);
    OffDiagScalar<NS, NT>(p.gas, sn, dun, fa, positive, od, extra, extraT);
  }
}

template <int NS, int NT, bool LOWER, bool UPPER, int JAC = kJacScalar>
__device__ __forceinline__ void OffDiagonals(const BlockDev &b, const Params &p,
                                             const double *__restrict__ x, int i, int j, int k,
                                             long long idx, double *L, double *U) {
  using E = Eq<NS, NT>;
  double own[E::neq];
  if (JAC == kJacRoe) LoadCell<E::neq>(b.state, b.fs, idx, own);
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
        OffDiagOne<NS, NT, JAC>(b, p, sn, dun, own, fa, true, idx - st,
                                p.isViscous ? __ldg(b.dist[d] + idx) : 1.0, od);
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
        OffDiagOne<NS, NT, JAC>(b, p, sn, dun, own, fa, false, idx + st,
                                p.isViscous ? __ldg(b.dist[d] + idx + st) : 1.0, od);
#pragma unroll
        for (int e = 0; e < E::neq; ++e) U[e] += od[e];
      }
    }
  }
}

// out = M v with M the cell's diagonal (or inverse) in `field`: {flow, turbulence} scalars, or
// the flow block fs x fs followed by the turbulence block NT x NT
// (ArrayMultiplication, ref: include/fluxJacobian.hpp:50-88)
template <int NS, int NT, int JAC>
__device__ __forceinline__ void DiagMult(const double *__restrict__ field, long long fs,
                                         long long idx, const double *v, double *out) {
  using E = Eq<NS, NT>;
  if (JAC == kJacBlock) {
    constexpr int nf = Blk<NS, NT>::fs;
#pragma unroll
    for (int rr = 0; rr < nf; ++rr) {
      double acc = 0.0;
#pragma unroll
      for (int cc = 0; cc < nf; ++cc) acc += __ldg(field + (rr * nf + cc) * fs + idx) * v[cc];
      out[rr] = acc;
    }
#pragma unroll
    for (int rr = 0; rr < NT; ++rr) {
      double acc = 0.0;
#pragma unroll
      for (int cc = 0; cc < NT; ++cc)
        acc += __ldg(field + (nf * nf + rr * NT + cc) * fs + idx) * v[nf + cc];
      out[nf + rr] = acc;
    }
  } else {
    const double dF = __ldg(field + idx);
    const double dT = NT > 0 ? __ldg(field + fs + idx) : 0.0;
#pragma unroll
    for (int e = 0; e < E::neq; ++e) out[e] = v[e] * (e < NS + 4 ? dF : dT);
  }
}

// K6 DPLUR (Jacobi) sweep: xout = D^-1 (b + L(xin) - U(xin)); ref: src/linearSolver.cpp:473-507
template <int NS, int NT, int JAC = kJacScalar>
__global__ void __launch_bounds__(JAC == kJacScalar ? 256 : 128)
    DplurKernel(BlockDev b, Params p, const double *__restrict__ xin, double *__restrict__ xout) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double L[E::neq], U[E::neq], rhs[E::neq], xn[E::neq];
  OffDiagonals<NS, NT, true, true, JAC>(b, p, xin, i, j, k, idx, L, U);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) rhs[e] = (__ldg(b.rhs + e * b.fs + idx) + 0.0) + (L[e] - U[e]);
  DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, rhs, xn);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) xout[e * b.fs + idx] = xn[e];
}

// K7 LU-SGS along one i+j+k hyperplane; ref: src/linearSolver.cpp:341-428. Cells of a plane do
// not couple, so any order inside the plane reproduces the reference's lexicographic result.
template <int NS, int NT, bool FORWARD, int JAC = kJacScalar>
__global__ void __launch_bounds__(128)
    LusgsPlaneKernel(BlockDev b, Params p, int plane, int fullGS) {
  using E = Eq<NS, NT>;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= b.nj || k >= b.nk) return;
  const int i = plane - j - k;
  if (i < 0 || i >= b.ni) return;
  const long long idx = CellIdx(b, i, j, k);
  double L[E::neq], U[E::neq], rhs[E::neq], xn[E::neq];
  if (FORWARD) {
    if (fullGS) {
      OffDiagonals<NS, NT, true, true, JAC>(b, p, b.x, i, j, k, idx, L, U);
    } else {
      OffDiagonals<NS, NT, true, false, JAC>(b, p, b.x, i, j, k, idx, L, U);
    }
#pragma unroll
    for (int e = 0; e < E::neq; ++e) rhs[e] = __ldg(b.rhs + e * b.fs + idx) + (L[e] - U[e]);
    DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, rhs, xn);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) b.x[e * b.fs + idx] = xn[e];
  } else {
    if (fullGS) {
      OffDiagonals<NS, NT, true, true, JAC>(b, p, b.x, i, j, k, idx, L, U);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) rhs[e] = (__ldg(b.rhs + e * b.fs + idx) + L[e]) - U[e];
      DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, rhs, xn);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) b.x[e * b.fs + idx] = xn[e];
    } else {
      OffDiagonals<NS, NT, false, true, JAC>(b, p, b.x, i, j, k, idx, L, U);
      DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, U, xn);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) {
        const double xo = b.x[e * b.fs + idx];
        b.x[e * b.fs + idx] = xo - xn[e];
      }
    }
  }
}

// K7, latency-split form: the plane kernels are bound by the dependent chain of ONE thread working
// through its up to six neighbours (~2 us each: gathers, U + dU -> primitives, two fluxes), while a
// hyperplane offers few cells. Here eight lanes share a cell -- lanes 0..2 take the lower
// neighbours in i, j, k, lanes 3..5 the upper ones -- and lane 0 collects the six products with
// shuffles in the reference's order (L: i, j, k; U: i, j, k; src/procBlock.cpp:1056-1170), so the
// sums are the same expressions as in LusgsPlaneKernel.
template <int NS, int NT, bool FORWARD, int JAC = kJacScalar>
__global__ void __launch_bounds__(256)
    LusgsPlaneSplitKernel(BlockDev b, Params p, int plane, int fullGS) {
  using E = Eq<NS, NT>;
  const int lane8 = threadIdx.x & 7;
  const int cell = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const int j = cell % b.nj, k = cell / b.nj;
  const int i = plane - j - k;
  const bool valid = k < b.nk && i >= 0 && i < b.ni;
  const long long idx = valid ? CellIdx(b, i, j, k) : 0;
  const bool doLower = FORWARD || fullGS != 0, doUpper = !FORWARD || fullGS != 0;
  double od[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) od[e] = 0.0;
  if (valid && lane8 < 6) {
    const int d = lane8 % 3;
    const bool upper = lane8 >= 3;
    const int c[3] = {i, j, k}, nd[3] = {b.ni, b.nj, b.nk};
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    const long long st = Stride(b, d);
    const bool wanted = upper ? doUpper : doLower;
    const bool contributes =
        upper ? (c[d] < nd[d] - 1 || ConnAcross(b, 2 * d + 2, c[d1], nd[d1], c[d2]))
              : (c[d] > 0 || ConnAcross(b, 2 * d + 1, c[d1], nd[d1], c[d2]));
    if (wanted && contributes) {
      const long long nidx = upper ? idx + st : idx - st;
      const long long fidx = upper ? idx + st : idx;  // the face between the two cells
      double sn[E::neq], dun[E::neq], own[E::neq], fa[4];
      LoadCell<E::neq>(b.state, b.fs, nidx, sn);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) dun[e] = b.x[e * b.fs + nidx];
      if (JAC == kJacRoe) LoadCell<E::neq>(b.state, b.fs, idx, own);
#pragma unroll
      for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[d] + q * b.fs + fidx);
      OffDiagOne<NS, NT, JAC>(b, p, sn, dun, own, fa, !upper, nidx,
                              p.isViscous ? __ldg(b.dist[d] + fidx) : 1.0, od);
    }
  }
  // lane 0 of the group: L = ((0 + od_i) + od_j) + od_k from lanes 0..2, U from lanes 3..5
  const int base = (threadIdx.x & 31) & ~7;
  double L[E::neq], U[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) {
    double l = 0.0, u = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      l += __shfl_sync(0xffffffffu, od[e], base + q);
      u += __shfl_sync(0xffffffffu, od[e], base + 3 + q);
    }
    L[e] = l;
    U[e] = u;
  }
  if (!valid || lane8 != 0) return;
  double rhs[E::neq], xn[E::neq];
  if (FORWARD) {
#pragma unroll
    for (int e = 0; e < E::neq; ++e) rhs[e] = __ldg(b.rhs + e * b.fs + idx) + (L[e] - U[e]);
    DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, rhs, xn);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) b.x[e * b.fs + idx] = xn[e];
  } else if (fullGS) {
#pragma unroll
    for (int e = 0; e < E::neq; ++e) rhs[e] = (__ldg(b.rhs + e * b.fs + idx) + L[e]) - U[e];
    DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, rhs, xn);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) b.x[e * b.fs + idx] = xn[e];
  } else {
    DiagMult<NS, NT, JAC>(b.dinv, b.fs, idx, U, xn);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double xo = b.x[e * b.fs + idx];
      b.x[e * b.fs + idx] = xo - xn[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// block-matrix diagonal (blusgs / bdplur).
// face state reconstructed from this cell's side towards its lower (UPPER_FACE = false: the
// face's "upper" state) or upper face (the face's "lower" state) -- the very expressions of
// FaceStates (march.cuh), one side only; ref: src/procBlock.cpp:399-431
template <int NS, int NT, int RECON, int LIM, bool UPPER_FACE>
__device__ __forceinline__ void OneSidedFaceState(const BlockDev &b, const Params &p, int d,
                                                  long long idx, double *out) {
  using E = Eq<NS, NT>;
  const long long st = Stride(b, d);
  if (RECON == AITHER_RECON_CONSTANT) {
    LoadCell<E::neq>(b.state, b.fs, idx, out);
  } else if (RECON == AITHER_RECON_MUSCL) {
    const double cLo = __ldg(b.mc[d] + idx), cHi = __ldg(b.mc[d] + b.fs + idx);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double um = __ldg(b.state + e * b.fs + idx - st), u0 = __ldg(b.state + e * b.fs + idx),
                   up = __ldg(b.state + e * b.fs + idx + st);
      out[e] = UPPER_FACE ? Muscl1<LIM>(um, u0, up, p.kappa, cHi, cLo)
                          : Muscl1<LIM>(up, u0, um, p.kappa, cLo, cHi);
    }
  } else {
    double w[5];
#pragma unroll
    for (int o = 0; o < 5; ++o)
      w[o] = __ldg(b.cw[d] + idx + (UPPER_FACE ? (o - 2) : (2 - o)) * st);
    const WenoGeom g = WenoSetup(w);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      double y[5];
#pragma unroll
      for (int o = 0; o < 5; ++o)
        y[o] = __ldg(b.state + e * b.fs + idx + (UPPER_FACE ? (o - 2) : (2 - o)) * st);
      out[e] = p.wenoZ ? Weno1<true>(g, y[0], y[1], y[2], y[3], y[4])
                       : Weno1<false>(g, y[0], y[1], y[2], y[3], y[4]);
    }
  }
}

// inviscid part of the block diagonal: for i, j, k: A -= dF_Ur(upper state of the lower face),
// A += dF_Ul(lower state of the upper face); ref: src/procBlock.cpp:447-486
template <int NS, int NT, int RECON, int LIM>
__global__ void __launch_bounds__(128) BlockDiagInvKernel(BlockDev b, Params p) {
  using B = Blk<NS, NT>;
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double A[B::n], J[B::n], fsd[E::neq], fa[4];
#pragma unroll
  for (int q = 0; q < B::n; ++q) A[q] = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const long long st = Stride(b, d);
    OneSidedFaceState<NS, NT, RECON, LIM, false>(b, p, d, idx, fsd);
#pragma unroll
    for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[d] + q * b.fs + idx);
    RusanovFluxJacobian<NS, NT>(p.gas, fsd, fa, false, J);
#pragma unroll
    for (int q = 0; q < B::n; ++q) A[q] -= J[q];
    OneSidedFaceState<NS, NT, RECON, LIM, true>(b, p, d, idx, fsd);
#pragma unroll
    for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[d] + q * b.fs + idx + st);
    RusanovFluxJacobian<NS, NT>(p.gas, fsd, fa, true, J);
#pragma unroll
    for (int q = 0; q < B::n; ++q) A[q] += J[q];
  }
#pragma unroll
  for (int q = 0; q < B::n; ++q) b.diag[q * b.fs + idx] = A[q];
}

// K5 for the block diagonal: time step; D <- relax on the diagonal entries + V(1+zeta)/(dt theta)
// [+ max(lambda)/CFL_dual]; D^-1 by the reference's Gauss-Jordan; b and x0 = D^-1 b
// (ref: src/linearSolver.cpp:111-188, include/matMultiArray3d.hpp:109-122)
template <int NS, int NT>
__global__ void __launch_bounds__(128)
    PrepBlockKernel(BlockDev b, Params p, double cfl, int bits, int *__restrict__ singular) {
  using B = Blk<NS, NT>;
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  const double vol = __ldg(b.vol + idx);
  const double srMax = fmax(b.specRad[idx], b.specRad[b.fs + idx]);
  double dt;
  if (bits & kPrepDt) {
    dt = p.dtNondim > 0.0 ? p.dtNondim : cfl * (vol / srMax);
    b.dt[idx] = dt;
  } else {
    dt = b.dt[idx];
  }
  if (bits & kPrepDiag) {
    double diagVolTime = (vol * (1.0 + p.zeta)) / (dt * p.theta);
    if (p.dualTimeCFL > 0.0) diagVolTime += srMax / p.dualTimeCFL;
    double A[B::n];
#pragma unroll
    for (int q = 0; q < B::n; ++q) A[q] = b.diag[q * b.fs + idx];
#pragma unroll
    for (int r = 0; r < B::fs; ++r) {
      A[r * B::fs + r] *= p.relax;
      A[r * B::fs + r] += diagVolTime;
    }
#pragma unroll
    for (int r = 0; r < NT; ++r) {
      A[B::nf + r * NT + r] *= p.relax;
      A[B::nf + r * NT + r] += diagVolTime;
    }
#pragma unroll
    for (int q = 0; q < B::n; ++q) b.diag[q * b.fs + idx] = A[q];
    bool ok = MatrixInverse<B::fs>(A);
    if (NT > 0) ok = MatrixInverse<(NT > 0 ? NT : 1)>(A + B::nf) && ok;
    if (!ok) *singular = 1;  // the reference exits here (src/matrix.cpp:83-86)
#pragma unroll
    for (int q = 0; q < B::n; ++q) b.dinv[q * b.fs + idx] = A[q];
  }
  if (bits & kPrepInit) {
    double s[E::neq], rb[E::neq], x0[E::neq];
    LoadCell<E::neq>(b.state, b.fs, idx, s);
    RhsB<NS, NT>(b, p, idx, s, vol, dt, rb);
    StoreCell<E::neq>(b.rhs, b.fs, idx, rb);
    DiagMult<NS, NT, kJacBlock>(b.dinv, b.fs, idx, rb, x0);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) b.x[e * b.fs + idx] = p.matrixRequiresInit ? x0[e] : 0.0;
  }
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
template <int NS, int NT, int JAC = kJacScalar>
__global__ void __launch_bounds__(JAC == kJacScalar ? 256 : 128)
    AxmbKernel(BlockDev b, Params p, double *__restrict__ partials, int storeField) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  double sq = 0.0;
  if (i < b.ni && j < b.nj) {
    const long long idx = CellIdx(b, i, j, k);
    double L[E::neq], U[E::neq], xc[E::neq], ax[E::neq];
    OffDiagonals<NS, NT, true, true, JAC>(b, p, b.x, i, j, k, idx, L, U);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) xc[e] = b.x[e * b.fs + idx];
    DiagMult<NS, NT, JAC>(b.diag, b.fs, idx, xc, ax);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double rb = __ldg(b.rhs + e * b.fs + idx);
      const double mr = 0.0 - ((ax[e] - (L[e] - U[e])) - rb);
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

// arg-max over a thread block, one candidate per block (deterministic: ties by traversal key)
__device__ __forceinline__ void BlockLinfToPartials(LinfCand best, LinfCand *linfPartials,
                                                    int blockLinear, int tid, int nthreads) {
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

constexpr int kUpdPlanes = 4;  // k-planes per thread block: the block reductions are paid once
// MOVE = false: the state has already been advanced by the matrix-residual pass
// (implicit_tma.cuh: every cell's updated primitive state is in registers there and is written to
// the alternate state buffer); what is left is the residual norms
template <int NS, int NT, bool NORMS = true, bool MOVE = true>
__global__ void __launch_bounds__(256)
    UpdateKernel(BlockDev b, Params p, double *__restrict__ partials,
                 LinfCand *__restrict__ linfPartials) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  double sq[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) sq[e] = 0.0;
  LinfCand best;
  best.v = 0.0;  // the reference starts from linf = 0 and uses a strict '>' (resid.hpp:33)
  best.key = 0x7fffffffffffffffLL;
  if (i < b.ni && j < b.nj) {
    const int kEnd = min(b.nk, (static_cast<int>(blockIdx.z) + 1) * kUpdPlanes);
    for (int k = blockIdx.z * kUpdPlanes; k < kEnd; ++k) {
      const long long idx = CellIdx(b, i, j, k);
      if (MOVE) {
        double s[E::neq], du[E::neq], sn[E::neq];
        LoadCell<E::neq>(b.state, b.fs, idx, s);
        LoadCell<E::neq>(b.x, b.fs, idx, du);
        UpdatePrimWithCons<NS, NT>(p.gas, s, du, sn);
        StoreCell<E::neq>(b.state, b.fs, idx, sn);
      }
      if (NORMS) {
        const long long cellKey =
            ((static_cast<long long>(k) * b.nj + j) * b.ni + i) * static_cast<long long>(E::neq);
#pragma unroll
        for (int e = 0; e < E::neq; ++e) {
          const double r = __ldg(b.resid + e * b.fs + idx);
          sq[e] += r * r;
          LinfCand c;
          c.v = r;
          c.key = cellKey + e;
          if (r > best.v) best = c;  // increasing k, e: first maximum wins, like the reference loop
        }
      }
    }
  }
  if (!NORMS) return;
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  const int nthreads = blockDim.x * blockDim.y;
  const int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  BlockSumToPartials<E::neq>(sq, partials, blockLinear, tid, nthreads);
  BlockLinfToPartials(best, linfPartials, blockLinear, tid, nthreads);
}

// final pass over the per-block partials of one procBlock, accumulating into the iteration's
// result record in a fixed order (deterministic run to run)
constexpr int kFinalThreads = 512;
static __global__ void __launch_bounds__(kFinalThreads)
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
static __global__ void __launch_bounds__(kFinalThreads)
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

// ---------------------------------------------------------------------------------------------
// output staging: one function-file variable for the physical cells of a block, i fastest
// (ref: WriteFunFile, src/output.cpp:229-330 -- each branch cited by its variable name there)
template <int NS, int NT>
__global__ void __launch_bounds__(256)
    OutputVarKernel(BlockDev b, Params p, int var, int species, double scale,
                    double *__restrict__ dst) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double s[E::neq];
  LoadCell<E::neq>(b.state, b.fs, idx, s);
  double v = 0.0;
  switch (var) {
    case AITHER_OUT_DENSITY: v = SpeciesSum<NS>(s); break;
    case AITHER_OUT_VEL_X: v = s[E::imx]; break;
    case AITHER_OUT_VEL_Y: v = s[E::imy]; break;
    case AITHER_OUT_VEL_Z: v = s[E::imz]; break;
    case AITHER_OUT_PRESSURE: v = s[E::ie]; break;
    case AITHER_OUT_MACH: v = sqrt(VelMagSq<NS>(s)) / SoS<NS>(p.gas, s); break;
    case AITHER_OUT_SOS: v = SoS<NS>(p.gas, s); break;
    case AITHER_OUT_DT: v = b.dt[idx]; break;
    // the reference writes its stored temperature field (refreshed from the state every residual
    // evaluation); for the current state that is T = p / sum(rho_s R_s)
    case AITHER_OUT_TEMPERATURE: v = Temperature<NS>(p.gas, s); break;
    case AITHER_OUT_ENERGY: v = Energy<NS>(p.gas, s); break;
    case AITHER_OUT_ENTHALPY: v = Enthalpy<NS>(p.gas, s); break;
    case AITHER_OUT_CP: v = Mixture<NS>(p.gas, s).cp; break;
    case AITHER_OUT_CV: v = Mixture<NS>(p.gas, s).cv; break;
    case AITHER_OUT_VISCOSITY_RATIO:
      v = (NT > 0 && b.eddyVisc) ? b.eddyVisc[idx] / b.viscosity[idx] : 0.0;
      break;
    case AITHER_OUT_TURBULENT_VISCOSITY: v = b.eddyVisc ? b.eddyVisc[idx] : 0.0; break;
    case AITHER_OUT_VISCOSITY: v = b.viscosity ? b.viscosity[idx] : 0.0; break;
    case AITHER_OUT_TKE: v = NT > 0 ? s[E::it] : 0.0; break;
    case AITHER_OUT_SDR: v = NT > 1 ? s[E::it + (NT > 1 ? 1 : 0)] : 0.0; break;
    case AITHER_OUT_F1: v = b.f1 ? b.f1[idx] : 0.0; break;
    case AITHER_OUT_F2: v = b.f2 ? b.f2[idx] : 0.0; break;
    case AITHER_OUT_WALL_DISTANCE: v = b.wallDist ? b.wallDist[idx] : 0.0; break;
    case AITHER_OUT_MASS_FRACTION: v = s[species < NS ? species : 0] / SpeciesSum<NS>(s); break;
    default: break;
  }
  dst[(static_cast<long long>(k) * b.nj + j) * b.ni + i] = v * scale;
}

static __global__ void FillKernel(double *__restrict__ p, long long n, double v) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x)
    p[t] = v;
}

}  // namespace aither
