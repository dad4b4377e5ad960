// march.cuh -- plane-marching kernels for the two hot loops of the iteration.
//
// Both kernels give one thread block a 32 x 8 (i x j) column of cells and march it along k
// through a chunk of planes, keeping what neighbouring cells need in shared memory:
//
//   ResidualMarchKernel  (K1: procBlock::CalcInvFluxI/J/K, src/procBlock.cpp:384,522,660)
//       ring of 2H+1 state planes (H = stencil half width) filled with cp.async one plane ahead;
//       every face flux is computed once (tile-edge faces twice) and handed to its two owner
//       cells through shared memory (i, j) or a register carried to the next plane (k): owner
//       writes, no atomics, the reference's accumulation order per cell.
//   ImplicitMarchKernel  (K6 dplur::DPLUR + K8 linearSolver::AXmB, src/linearSolver.cpp:473,58)
//       each cell's "off-diagonal ingredients" (state, U + dU converted back to primitives,
//       enthalpies, speed of sound) are computed ONCE per sweep by the owning thread and shared
//       through shared memory, instead of six times (once per neighbour) as in the reference's
//       ImplicitLower/Upper (src/procBlock.cpp:1056-1170). k-neighbours are carried in registers.
//
// The fp64 pipe, not HBM, is what the reference formulas saturate first on B200 (ncu:
// profiles/r01a_*): these kernels therefore also restructure the point-wise maths (one reciprocal
// shared by several quotients, |v|^2 without the sqrt round trip, thermodynamic sums collapsed for
// a single species). Every change is a rounding-level (<= few ulp) reassociation of the reference
// formula, cited in place; parity is held to 1e-12 by tests/test_gpu_*.py.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace aither {

constexpr int kMI = 32, kMJ = 8, kMThreads = kMI * kMJ;

// MUSCL with the two grid-ratio coefficients precomputed per cell (MusclCoefKernel):
// dPlus = 2 w1 / (w1 + wd), dMinus = 2 w1 / (w1 + w2); ref include/reconstruction.hpp:128-153
template <int NEQ, int LIM>
__device__ __forceinline__ void MusclC(const double *u2, const double *u1, const double *d1,
                                       double kappa, double dPlus, double dMinus, double *face) {
#pragma unroll
  for (int e = 0; e < NEQ; ++e) face[e] = Muscl1<LIM>(u2[e], u1[e], d1[e], kappa, dPlus, dMinus);
}

// grid-ratio coefficients of every cell along direction d:
//   mc[0] = (w + w) / (w + w_lower)   mc[1] = (w + w) / (w + w_upper)
// (the exact expressions of reconstruction.hpp:133-134, so the values are bit-identical to the
// ones the reference forms per face)
static __global__ void MusclCoefKernel(BlockDev b, int d) {
  const int NI = b.ni + 2 * b.g, NJ = b.nj + 2 * b.g, NK = b.nk + 2 * b.g;
  const long long n = static_cast<long long>(NI) * NJ * NK;
  const int nd[3] = {NI, NJ, NK};
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c[3] = {static_cast<int>(t % NI), static_cast<int>((t / NI) % NJ),
                      static_cast<int>(t / (static_cast<long long>(NI) * NJ))};
    const long long idx = CellIdx(b, c[0] - b.g, c[1] - b.g, c[2] - b.g);
    const long long st = Stride(b, d);
    const double w = b.cw[d][idx];
    const double lo = c[d] > 0 ? (w + w) / (w + b.cw[d][idx - st]) : 0.0;
    const double hi = c[d] < nd[d] - 1 ? (w + w) / (w + b.cw[d][idx + st]) : 0.0;
    b.mc[d][idx] = lo;
    b.mc[d][b.fs + idx] = hi;
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void CpAsync8(double *smemDst, const double *gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smemDst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_all;\n" ::); }

template <int RECON>
struct Halo {
  static constexpr int H = RECON == AITHER_RECON_CONSTANT ? 1 : (RECON == AITHER_RECON_MUSCL ? 2 : 3);
};

template <int NS, int NT, int RECON>
struct ResSmem {
  static constexpr int H = Halo<RECON>::H;
  static constexpr int PI = kMI + 2 * H, PJ = kMJ + 2 * H, PC = PI * PJ;
  static constexpr int NSLOT = 2 * H + 1;
  static constexpr int neq = NS + 4 + NT;
  static constexpr int FI = (kMI + 1) * kMJ, FJ = kMI * (kMJ + 1);
  static constexpr size_t bytes = sizeof(double) * (static_cast<size_t>(NSLOT) * neq * PC +
                                                    static_cast<size_t>(neq) * (FI + FJ));
};

// Reconstruct the two face states from 2H stencil cells. `ld(o, e)` returns component e of the
// cell at offset o in [-H, H-1] from the face (o = -1: lower cell, o = 0: upper cell).
template <int NS, int NT, int RECON, int LIM, typename LD>
__device__ __forceinline__ void FaceStates(const BlockDev &b, const Params &p, int d,
                                           long long idx, LD ld, double *fl, double *fr) {
  using E = Eq<NS, NT>;
  if (RECON == AITHER_RECON_CONSTANT) {
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      fl[e] = ld(-1, e);
      fr[e] = ld(0, e);
    }
  } else if (RECON == AITHER_RECON_MUSCL) {
    const long long st = Stride(b, d);
    const double *mc = b.mc[d];
    // lower cell (idx - st): dPlus = its upper ratio, dMinus = its lower ratio; upper cell (idx):
    // dPlus = its lower ratio, dMinus = its upper ratio (see FaceFlux in kernels.cuh)
    const double lLo = __ldg(mc + idx - st), lHi = __ldg(mc + b.fs + idx - st);
    const double uLo = __ldg(mc + idx), uHi = __ldg(mc + b.fs + idx);
#pragma unroll
    for (int e = 0; e < E::neq; ++e) {
      const double um2 = ld(-2, e), um1 = ld(-1, e), u0 = ld(0, e), up1 = ld(1, e);
      fl[e] = Muscl1<LIM>(um2, um1, u0, p.kappa, lHi, lLo);
      fr[e] = Muscl1<LIM>(up1, u0, um1, p.kappa, uLo, uHi);
    }
  } else {
    const long long st = Stride(b, d);
    double w[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) w[o] = __ldg(b.cw[d] + idx + (o - 3) * st);
    {
      const double wl[5] = {w[0], w[1], w[2], w[3], w[4]};
      const WenoGeom g = WenoSetup(wl);
#pragma unroll
      for (int e = 0; e < E::neq; ++e)
        fl[e] = p.wenoZ ? Weno1<true>(g, ld(-3, e), ld(-2, e), ld(-1, e), ld(0, e), ld(1, e))
                        : Weno1<false>(g, ld(-3, e), ld(-2, e), ld(-1, e), ld(0, e), ld(1, e));
    }
    {
      const double wr[5] = {w[5], w[4], w[3], w[2], w[1]};
      const WenoGeom g = WenoSetup(wr);
#pragma unroll
      for (int e = 0; e < E::neq; ++e)
        fr[e] = p.wenoZ ? Weno1<true>(g, ld(2, e), ld(1, e), ld(0, e), ld(-1, e), ld(-2, e))
                        : Weno1<false>(g, ld(2, e), ld(1, e), ld(0, e), ld(-1, e), ld(-2, e));
    }
  }
}

// One face: reconstruct from the stencil cells at shared-memory offsets off[0..2H) (component
// stride PC), Riemann flux, times area. `gidx` = global index of the cell on the upper side.
template <int NS, int NT, int RECON, int LIM, int FLUX, int PC>
__device__ __forceinline__ void FaceFluxSmem(const BlockDev &b, const Params &p, int d,
                                             long long gidx, const double *sm, const int *off,
                                             double *out) {
  using E = Eq<NS, NT>;
  constexpr int H = Halo<RECON>::H;
  const double *fa = d == 0 ? b.fA[0] : (d == 1 ? b.fA[1] : b.fA[2]);
  double area[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) area[q] = __ldg(fa + q * b.fs + gidx);
  auto ld = [&](int o, int e) { return sm[e * PC + off[o + H]]; };
  double fl[E::neq], fr[E::neq];
  FaceStates<NS, NT, RECON, LIM>(b, p, d, gidx, ld, fl, fr);
  InviscidFluxFast<NS, NT, FLUX>(p.gas, fl, fr, area, out);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) out[e] *= area[3];
}

template <int NS, int NT, int RECON, int LIM, int FLUX>
__global__ void __launch_bounds__(kMThreads, 2)
    ResidualMarchKernel(BlockDev b, Params p, int kChunk, int implicitScalar, int fusePrep,
                        double cfl) {
  using E = Eq<NS, NT>;
  using S = ResSmem<NS, NT, RECON>;
  constexpr int H = S::H, PI = S::PI, PC = S::PC, NSLOT = S::NSLOT;
  constexpr int SLOTSZ = E::neq * PC;
  constexpr int FTOT = S::FI + S::FJ;
  extern __shared__ double smem[];
  double *ring = smem;                                      // [NSLOT][neq][PC]
  double *sfl = smem + static_cast<size_t>(NSLOT) * SLOTSZ;  // [neq][FI + FJ]

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + kMI * ty;
  const int i0 = blockIdx.x * kMI, j0 = blockIdx.y * kMJ;
  const int k0 = blockIdx.z * kChunk;
  const int k1 = min(k0 + kChunk, b.nk);
  const int i = i0 + tx, j = j0 + ty;
  const bool colValid = i < b.ni && j < b.nj;
  const int iMaxPad = b.ni + b.g - 1, jMaxPad = b.nj + b.g - 1, kMaxPad = b.nk + b.g - 1;

  auto slotOf = [&](int kk) { return ((kk % NSLOT) + NSLOT) % NSLOT; };
  auto loadPlane = [&](int kk) {
    // plane kk (may be a ghost plane) with an in-plane halo of H, clamped to the padded block
    double *dst = ring + slotOf(kk) * SLOTSZ;
    const int kc = min(max(kk, -b.g), kMaxPad);
    for (int c = tid; c < PC; c += kMThreads) {
      const int pi = c % PI, pj = c / PI;
      const int gi = min(max(i0 - H + pi, -b.g), iMaxPad);
      const int gj = min(max(j0 - H + pj, -b.g), jMaxPad);
      const long long gidx = CellIdx(b, gi, gj, kc);
#pragma unroll
      for (int e = 0; e < E::neq; ++e) CpAsync8(dst + e * PC + c, b.state + e * b.fs + gidx);
    }
  };

  // prologue: planes k0-H .. k0+H-1
  for (int kk = k0 - H; kk < k0 + H; ++kk) loadPlane(kk);
  CpAsyncCommit();

  const int pc = (tx + H) + PI * (ty + H);  // this thread's cell inside a plane
  // flux slots of this thread's cell: lower / upper i-face, lower / upper j-face
  const int fIlo = tx + (kMI + 1) * ty, fJlo = S::FI + tx + kMI * ty;
  double pend[E::neq];                       // residual of the cell one plane below, k-hi missing
  double pendSpec = 0.0, pendSpecT = 0.0, sosPrev = 0.0;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) pend[e] = 0.0;

  for (int k = k0; k <= k1; ++k) {
    CpAsyncWaitAll();
    __syncthreads();  // planes k-H..k+H-1 have landed; the previous step's readers are done
    if (k < k1) loadPlane(k + H);  // prefetch for the next step (index clamped inside)
    CpAsyncCommit();

    const long long idx = CellIdx(b, i, j, k);
    // geometry of the next plane comes through registers (no in-plane reuse): pull its lines
    // into L2 one plane ahead, one request per 128-byte line (lanes 0 and 16 of a row)
    if (p.prefetch && (tx & 15) == 0 && colValid && k + p.prefetch <= b.nk) {
      const long long idxn = idx + p.prefetch * b.sk;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) PrefetchL2(b.fA[d] + q * b.fs + idxn);
        if (RECON == AITHER_RECON_MUSCL) {
          PrefetchL2(b.mc[d] + idxn);
          PrefetchL2(b.mc[d] + b.fs + idxn);
        } else if (RECON != AITHER_RECON_CONSTANT) {
          PrefetchL2(b.cw[d] + idxn);
        }
      }
      if (fusePrep) PrefetchL2(b.vol + idxn);
    }
    const int curBase = slotOf(k) * SLOTSZ;
    double fk[E::neq];
#pragma unroll
    for (int e = 0; e < E::neq; ++e) fk[e] = 0.0;
    // face tasks of this thread, one code copy: 0 = k-face below the cell, 1 = i-face, 2 = j-face,
    // 3 = tile-edge faces (i-faces at lx = 32: warp 0 lanes 0..7; j-faces at ly = 8: warp 1)
    const int nTask = k == k1 ? 1 : 4;
#pragma unroll 1
    for (int t = 0; t < nTask; ++t) {
      int d, lx = tx, ly = ty, dst = -1;
      bool valid;
      if (t == 0) {
        d = 2;
        valid = colValid;
      } else if (t == 1) {
        d = 0;
        valid = i <= b.ni && j < b.nj;
        dst = fIlo;
      } else if (t == 2) {
        d = 1;
        valid = i < b.ni && j <= b.nj;
        dst = fJlo;
      } else if (tid < kMJ) {
        d = 0;
        lx = kMI;
        ly = tid;
        valid = i0 + kMI <= b.ni && j0 + ly < b.nj;
        dst = kMI + (kMI + 1) * ly;
      } else {
        d = 1;
        lx = tid - 32;
        ly = kMJ;
        valid = tid >= 32 && tid < 32 + kMI && i0 + lx < b.ni && j0 + kMJ <= b.nj;
        dst = S::FI + lx + kMI * kMJ;
      }
      if (!valid) continue;
      int off[2 * H];
      const int c0 = (lx + H) + PI * (ly + H);
      if (d == 2) {
#pragma unroll
        for (int o = 0; o < 2 * H; ++o) off[o] = slotOf(k + o - H) * SLOTSZ + c0;
      } else {
        const int st = d == 0 ? 1 : PI;
#pragma unroll
        for (int o = 0; o < 2 * H; ++o) off[o] = curBase + c0 + (o - H) * st;
      }
      double f[E::neq];
      FaceFluxSmem<NS, NT, RECON, LIM, FLUX, PC>(b, p, d, CellIdx(b, i0 + lx, j0 + ly, k), ring,
                                                 off, f);
      if (t == 0) {
#pragma unroll
        for (int e = 0; e < E::neq; ++e) fk[e] = f[e];
      } else {
#pragma unroll
        for (int e = 0; e < E::neq; ++e) sfl[e * FTOT + dst] = f[e];
      }
    }
    double fAk[4] = {0.0, 0.0, 0.0, 1.0};
    if (colValid) {
#pragma unroll
      for (int q = 0; q < 4; ++q) fAk[q] = __ldg(b.fA[2] + q * b.fs + idx);
    }
    // finalise the cell below (k-1): add its upper k-face flux, k-direction spectral radius
    if (colValid && k > k0) {
      const long long idxm = idx - b.sk;
      // the pending cell's state is still in the ring (plane k-1) and its lower k-face area in
      // L1: re-read both instead of carrying nine doubles through the face loop
      double sPrev[E::neq], fAkLo[4];
      const int prevBase = slotOf(k - 1) * SLOTSZ;
#pragma unroll
      for (int e = 0; e < E::neq; ++e) sPrev[e] = ring[prevBase + e * PC + pc];
#pragma unroll
      for (int q = 0; q < 4; ++q) fAkLo[q] = __ldg(b.fA[2] + q * b.fs + idxm);
      double res[E::neq];
#pragma unroll
      for (int e = 0; e < E::neq; ++e) {
        res[e] = pend[e] + fk[e];
        b.resid[e * b.fs + idxm] = res[e];
      }
      double srTk = 0.0;
      const double sr = pendSpec + InvCellSpectralRadii<NS>(sPrev, sosPrev, fAkLo, fAk, &srTk);
      const double srT = NT > 0 ? pendSpecT + srTk : 0.0;
      b.specRad[idxm] = sr;
      b.specRad[b.fs + idxm] = srT;
      if (!fusePrep) {
        if (implicitScalar) {
          b.diag[idxm] = sr;
          if (NT > 0) b.diag[b.fs + idxm] = srT;
        }
      } else {
        // the cell's residual and spectral radius are final here, so the time step, the scalar
        // diagonal and its inverse, the right-hand side and x0 = D^-1 b follow in the same pass
        // (PrepKernel's arithmetic, kernels.cuh; ref src/procBlock.cpp:782-821,
        // src/linearSolver.cpp:111-188) instead of re-reading residual and spectral radius
        const double vol = __ldg(b.vol + idxm);
        const double srMax = fmax(sr, 0.0);
        const double dt = p.dtNondim > 0.0 ? p.dtNondim : cfl * (vol / srMax);
        b.dt[idxm] = dt;
        double diagVolTime = (vol * (1.0 + p.zeta)) / (dt * p.theta);
        if (p.dualTimeCFL > 0.0) diagVolTime += srMax / p.dualTimeCFL;
        double a = sr;
        a *= p.relax;
        a += diagVolTime;
        b.diag[idxm] = a;
        const double dinv = 1.0 / a;
        b.dinv[idxm] = dinv;
        double rb[E::neq];
        RhsB<NS, NT, true>(b, p, idxm, sPrev, vol, dt, rb, res);
#pragma unroll
        for (int e = 0; e < E::neq; ++e) {
          b.rhs[e * b.fs + idxm] = rb[e];
          b.x[e * b.fs + idxm] = p.matrixRequiresInit ? rb[e] * dinv : 0.0;
        }
      }
    }
    if (k == k1) break;
    __syncthreads();
    if (colValid) {
      // reference accumulation order (src/procBlock.cpp:447-463): i-lo, i-hi, j-lo, j-hi, k-lo,
      // (k-hi at the next plane)
      double s[E::neq];
#pragma unroll
      for (int e = 0; e < E::neq; ++e) s[e] = ring[curBase + e * PC + pc];
      const MixK<NS> m = MixOf<NS>(p.gas, s);
      const double sos = sqrt(m.gamma * s[E::ie] * m.rhoInv);
      double aLo[4], aHi[4];
#pragma unroll
      for (int e = 0; e < E::neq; ++e) {
        double r = 0.0;
        r -= sfl[e * FTOT + fIlo];
        r += sfl[e * FTOT + fIlo + 1];
        r -= sfl[e * FTOT + fJlo];
        r += sfl[e * FTOT + fJlo + kMI];
        r -= fk[e];
        pend[e] = r;
      }
      double sr = 0.0, srT = 0.0, st1 = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        aLo[q] = __ldg(b.fA[0] + q * b.fs + idx);
        aHi[q] = __ldg(b.fA[0] + q * b.fs + idx + 1);
      }
      sr += InvCellSpectralRadii<NS>(s, sos, aLo, aHi, &st1);
      srT += st1;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        aLo[q] = __ldg(b.fA[1] + q * b.fs + idx);
        aHi[q] = __ldg(b.fA[1] + q * b.fs + idx + b.sj);
      }
      sr += InvCellSpectralRadii<NS>(s, sos, aLo, aHi, &st1);
      srT += st1;
      pendSpec = sr;
      pendSpecT = srT;
      sosPrev = sos;
    }
  }
  CpAsyncWaitAll();
}

// ---------------------------------------------------------------------------------------------
// implicit sweep (cell ingredients: MakeIngr / OffDiagFromIngr in physics.cuh)
constexpr int kIPI = kMI + 2, kIPJ = kMJ + 2, kIPC = kIPI * kIPJ;

enum ImplicitMode { kModeDplur = 0, kModeAxmb = 1 };

// MODE kModeDplur: xout = D^-1 (b + L(xin) - U(xin))      (ref src/linearSolver.cpp:473-507)
// MODE kModeAxmb : mr = -((D x - (L - U)) - b), partial sums of mr^2 per block (:58-109)
template <int NS, int NT, int MODE>
__global__ void __launch_bounds__(kMThreads, 2)
    ImplicitMarchKernel(BlockDev b, Params p, const double *__restrict__ xin,
                        double *__restrict__ xout, int kChunk, double *__restrict__ partials,
                        int storeField) {
  using E = Eq<NS, NT>;
  using G = Ingr<NS, NT>;
  constexpr int neq = E::neq;
  extern __shared__ double smem[];  // [2][G::n][kIPC]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + kMI * ty;
  const int i0 = blockIdx.x * kMI, j0 = blockIdx.y * kMJ;
  const int k0 = blockIdx.z * kChunk;
  const int k1 = min(k0 + kChunk, b.nk);
  const int i = i0 + tx, j = j0 + ty;
  const bool colValid = i < b.ni && j < b.nj;
  const int iMaxPad = b.ni + b.g - 1, jMaxPad = b.nj + b.g - 1;
  const int pc = (tx + 1) + kIPI * (ty + 1);

  // halo cell this thread also prepares (threads 0..83 cover the ring around the tile)
  int hpi = -1, hpj = -1;
  if (tid < 2 * kMI) {           // rows above / below the tile
    hpi = 1 + (tid % kMI);
    hpj = tid < kMI ? 0 : kIPJ - 1;
  } else if (tid < 2 * kMI + 2 * kMJ) {  // columns left / right
    const int q = tid - 2 * kMI;
    hpi = q < kMJ ? 0 : kIPI - 1;
    hpj = 1 + (q % kMJ);
  }
  const int hgi = min(max(i0 - 1 + hpi, -b.g), iMaxPad);
  const int hgj = min(max(j0 - 1 + hpj, -b.g), jMaxPad);

  double accLp[neq], accUp[neq];  // pending cell (k-1): complete L, U without the k+1 term
  double carryL[neq];             // L-term for this plane's cell, produced one plane below
  double sq = 0.0;
#pragma unroll
  for (int e = 0; e < neq; ++e) {
    accLp[e] = 0.0;
    accUp[e] = 0.0;
    carryL[e] = 0.0;
  }
  // ghost columns next to a partial tile are prepared too (their state is valid ghost data)
  const bool ingValid = i <= iMaxPad && j <= jMaxPad;

  for (int k = k0 - 1; k <= k1; ++k) {
    double *cur = smem + static_cast<size_t>((k - k0 + 1) & 1) * G::n * kIPC;
    const long long idx = CellIdx(b, i, j, k);
    const bool planeInterior = k >= k0 && k < k1;
    // everything this kernel reads comes through registers: pull the next plane's lines into L2
    // (one request per 128-byte line)
    if (p.prefetch && (tx & 15) == 0 && colValid && k + 1 <= b.nk) {
      const long long idxn = idx + b.sk;
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        PrefetchL2(b.state + e * b.fs + idxn);
        PrefetchL2(xin + e * b.fs + idxn);
        PrefetchL2(b.rhs + e * b.fs + idx);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) PrefetchL2(b.fA[d] + q * b.fs + idxn);
        if (p.isViscous) PrefetchL2(b.dist[d] + idxn);
      }
      if (p.isViscous) PrefetchL2(b.viscosity + idxn);
      PrefetchL2((MODE == kModeDplur ? b.dinv : b.diag) + idx);
    }
    double newCarry[neq];
#pragma unroll
    for (int e = 0; e < neq; ++e) newCarry[e] = 0.0;
    // ---- ingredients of this thread's cell in plane k; everything that needs only them ----
    if (ingValid) {
      double ing[G::n];
      {
        double s[neq], du[neq];
        LoadCell<neq>(b.state, b.fs, idx, s);
        LoadCell<neq>(xin, b.fs, idx, du);
        MakeIngr<NS, NT>(p.gas, s, du, &ing[neq], &ing[neq + 1], &ing[2 * neq + 2],
                         &ing[3 * neq + 2]);
#pragma unroll
        for (int e = 0; e < neq; ++e) {
          ing[e] = s[e];
          ing[neq + 2 + e] = du[e];
        }
        ing[G::ivt] = p.isViscous ? ViscSpecFactor(p.tr, SpeciesSum<NS>(s), Gamma<NS>(p.gas, s),
                                                   __ldg(b.viscosity + idx))
                                  : 0.0;
      }
      if (planeInterior) {
#pragma unroll
        for (int q = 0; q < G::n; ++q) cur[q * kIPC + pc] = ing[q];
      }
      auto ldOwn = [&](int q) { return ing[q]; };
      if (colValid && k > k0) {
        // U-term of the cell below (k-1) across face k, then finish that cell
        const bool useKhi = k < b.nk || ConnAcross(b, 6, i, b.ni, j);
        if (useKhi) {
          double fa[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[2] + q * b.fs + idx);
          OffDiagFromIngr<NS, NT>(ldOwn, fa, false, accUp,
                                  p.isViscous ? fa[3] / __ldg(b.dist[2] + idx) * ing[G::ivt] : 0.0);
        }
        const long long idxm = idx - b.sk;
        if (MODE == kModeDplur) {
          const double dinv = __ldg(b.dinv + idxm);
#pragma unroll
          for (int e = 0; e < neq; ++e) {
            const double rb = __ldg(b.rhs + e * b.fs + idxm);
            xout[e * b.fs + idxm] = ((rb + 0.0) + (accLp[e] - accUp[e])) * dinv;
          }
        } else {
          const double a = __ldg(b.diag + idxm);
#pragma unroll
          for (int e = 0; e < neq; ++e) {
            const double rb = __ldg(b.rhs + e * b.fs + idxm);
            const double ax = __ldg(xin + e * b.fs + idxm) * a;
            const double mr = 0.0 - ((ax - (accLp[e] - accUp[e])) - rb);
            if (storeField) b.mres[e * b.fs + idxm] = mr;
            sq += mr * mr;
          }
        }
      }
      if (colValid && k + 1 < k1) {
        // L-term this cell contributes to the cell above (k+1), across face k+1
        double fa[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[2] + q * b.fs + idx + b.sk);
        OffDiagFromIngr<NS, NT>(ldOwn, fa, true, newCarry,
                                p.isViscous ? fa[3] / __ldg(b.dist[2] + idx + b.sk) * ing[G::ivt]
                                            : 0.0);
      }
    }
    if (planeInterior && hpi >= 0) {
      const long long hidx = CellIdx(b, hgi, hgj, k);
      double s[neq], du[neq], hing[G::n];
      LoadCell<neq>(b.state, b.fs, hidx, s);
      LoadCell<neq>(xin, b.fs, hidx, du);
      MakeIngr<NS, NT>(p.gas, s, du, &hing[neq], &hing[neq + 1], &hing[2 * neq + 2],
                       &hing[3 * neq + 2]);
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        hing[e] = s[e];
        hing[neq + 2 + e] = du[e];
      }
      hing[G::ivt] = p.isViscous ? ViscSpecFactor(p.tr, SpeciesSum<NS>(s), Gamma<NS>(p.gas, s),
                                                  __ldg(b.viscosity + hidx))
                                 : 0.0;
      const int hc = hpi + kIPI * hpj;
#pragma unroll
      for (int q = 0; q < G::n; ++q) cur[q * kIPC + hc] = hing[q];
    }
    __syncthreads();
    if (colValid && planeInterior) {
      // which neighbours contribute: physical, or across a connection boundary
      // (ref src/procBlock.cpp:1064,1115)
      const bool useIlo = i > 0 || ConnAcross(b, 1, j, b.nj, k);
      const bool useIhi = i < b.ni - 1 || ConnAcross(b, 2, j, b.nj, k);
      const bool useJlo = j > 0 || ConnAcross(b, 3, k, b.nk, i);
      const bool useJhi = j < b.nj - 1 || ConnAcross(b, 4, k, b.nk, i);
      const bool useKlo = k > 0 || ConnAcross(b, 5, i, b.ni, j);
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        accLp[e] = 0.0;
        accUp[e] = 0.0;
      }
      double fa[4];
      if (useIlo) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[0] + q * b.fs + idx);
        auto ld = [&](int q) { return cur[q * kIPC + pc - 1]; };
        OffDiagFromIngr<NS, NT>(ld, fa, true, accLp,
                                p.isViscous ? fa[3] / __ldg(b.dist[0] + idx) * ld(G::ivt) : 0.0);
      }
      if (useJlo) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[1] + q * b.fs + idx);
        auto ld = [&](int q) { return cur[q * kIPC + pc - kIPI]; };
        OffDiagFromIngr<NS, NT>(ld, fa, true, accLp,
                                p.isViscous ? fa[3] / __ldg(b.dist[1] + idx) * ld(G::ivt) : 0.0);
      }
      if (useKlo) {  // produced from the cell below at the previous plane
#pragma unroll
        for (int e = 0; e < neq; ++e) accLp[e] += carryL[e];
      }
      if (useIhi) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[0] + q * b.fs + idx + 1);
        auto ld = [&](int q) { return cur[q * kIPC + pc + 1]; };
        OffDiagFromIngr<NS, NT>(ld, fa, false, accUp,
                                p.isViscous ? fa[3] / __ldg(b.dist[0] + idx + 1) * ld(G::ivt) : 0.0);
      }
      if (useJhi) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[1] + q * b.fs + idx + b.sj);
        auto ld = [&](int q) { return cur[q * kIPC + pc + kIPI]; };
        OffDiagFromIngr<NS, NT>(ld, fa, false, accUp,
                                p.isViscous ? fa[3] / __ldg(b.dist[1] + idx + b.sj) * ld(G::ivt)
                                            : 0.0);
      }
    }
#pragma unroll
    for (int e = 0; e < neq; ++e) carryL[e] = newCarry[e];
  }
  if (MODE == kModeAxmb) {
    const int blockLinear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    BlockSumToPartials<1>(&sq, partials, blockLinear, tid, kMThreads);
  }
}

}  // namespace aither
