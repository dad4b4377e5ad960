// lusgs_pencil.cuh -- scalar-diagonal LU-SGS as a persistent pencil wavefront with shared per-cell
// ingredients: one launch per half sweep, one thread per grid line.
//
// ref: src/linearSolver.cpp:341-428 (lusgs::LUSGS_Forward / LUSGS_Backward), hyperplane order
// src/utility.cpp:377-398, off-diagonals src/procBlock.cpp:1056-1170,
// src/fluxJacobian.cpp:122-162 (RusanovScalarOffDiagonal).
//
// Why the per-hyperplane kernels (and the eight-lanes-per-cell wavefront, lusgs_wave.cuh) sit at
// 4 % of the HBM roofline: not launch latency alone -- every cell evaluates SIX complete
// off-diagonal products (U + dU -> primitives, two fluxes, IEEE divisions: ~4 500 fp64
// instructions per cell against ~350 in the DPLUR march), i.e. they are bound by the fp64 pipe.
// This kernel does the arithmetic the DPLUR march does:
//
//   * per cell and half sweep the update-dependent ingredients (updated primitive state, its
//     enthalpy: MakeIngrDyn) are formed ONCE for the old and once for the new update; what does
//     not depend on the update at all (state, H, a, viscous spectral factors, b, D^-1) is packed
//     once per iteration into an array-of-structs record per cell (WaveDynKernel), the face areas
//     towards the "behind" side once per block (WaveGeoKernel): a thread walking its grid line
//     reads whole 32-byte sectors it uses completely instead of 8 bytes of ~30 different lines;
//   * "sweep space": a backward sweep is the forward sweep of the mirrored block. A cell needs
//     the NEW update of its three neighbours behind it and the OLD update of the three ahead;
//   * the block is cut into pencils of TJ x TK lines; a thread owns one line and visits cell
//     I = q - jl - kl at local plane q. Per plane, phase A: the thread forms the old ingredients
//     of the cell it will solve NEXT plane and pushes that cell's contribution to its three
//     behind-neighbours (own line: a register; j, k: shared memory); phase B: it solves its
//     cell from the records its behind-neighbours left in shared memory one plane ago (gather,
//     own faces) and the three pushes, writes the update, forms the new ingredients and leaves
//     its record. The six products are summed in the reference's order (i, j, k);
//   * cells of the neighbouring pencils (and the block's ghost cells across connections) are
//     served by one extra warp of halo threads; pencils are ordered by tickets along
//     anti-diagonals and wait on the progress counters of the two pencils behind them
//     (lusgs_wave.cuh), so a predecessor is always running or done.
#pragma once
#include <cuda_runtime.h>

#include "lusgs_wave.cuh"

namespace aither {

template <int NS, int NT>
struct PencilRec {
  static constexpr int neq = NS + 4 + NT;
  // per-iteration record: s[neq] | H a vt vtT | b[neq] | dinv dinvT
  static constexpr int DN = 2 * neq + 6;
  static constexpr int iH = neq, iA = neq + 1, iVt = neq + 2, iVtT = neq + 3, iB = neq + 4,
                       iD = 2 * neq + 4;
  // per-block record: behind-side faces i, j, k {nx, ny, nz, |A|} | |A| / dist for i, j, k | pad
  static constexpr int GN = 16;
  // shared-memory record of a solved / foreign cell: s | H a vt vtT | du | sn | Hn
  static constexpr int RN = 3 * neq + 5;
};

// index of cell (i, j, k), one ghost layer included, in the array-of-structs workspaces
__host__ __device__ __forceinline__ long long WaveIdx(const BlockDev &b, int i, int j, int k) {
  return (static_cast<long long>(k + 1) * (b.nj + 2) + (j + 1)) * (b.ni + 2) + (i + 1);
}

// behind-side faces: lower faces for the forward sweep (geoLo), upper faces for the backward one
static __global__ void WaveGeoKernel(BlockDev b, int isViscous, double *__restrict__ geoLo,
                                     double *__restrict__ geoHi) {
  const int NI = b.ni + 2, NJ = b.nj + 2, NK = b.nk + 2;
  const long long n = static_cast<long long>(NI) * NJ * NK;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t % NI) - 1, j = static_cast<int>((t / NI) % NJ) - 1;
    const int k = static_cast<int>(t / (static_cast<long long>(NI) * NJ)) - 1;
    const long long idx = CellIdx(b, i, j, k);
    double lo[16], hi[16];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const long long st = Stride(b, d);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        lo[4 * d + q] = b.fA[d][q * b.fs + idx];
        hi[4 * d + q] = b.fA[d][q * b.fs + idx + st];
      }
      lo[12 + d] = isViscous ? lo[4 * d + 3] / b.dist[d][idx] : 0.0;
      hi[12 + d] = isViscous ? hi[4 * d + 3] / b.dist[d][idx + st] : 0.0;
    }
    lo[15] = hi[15] = 0.0;
    double2 *oLo = reinterpret_cast<double2 *>(geoLo + t * 16);
    double2 *oHi = reinterpret_cast<double2 *>(geoHi + t * 16);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      oLo[q] = make_double2(lo[2 * q], lo[2 * q + 1]);
      oHi[q] = make_double2(hi[2 * q], hi[2 * q + 1]);
    }
  }
}

template <int NS, int NT>
__global__ void __launch_bounds__(256) WaveDynKernel(BlockDev b, Params p, double *__restrict__ dyn) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT>;
  constexpr int neq = E::neq;
  const int NI = b.ni + 2, NJ = b.nj + 2, NK = b.nk + 2;
  const long long n = static_cast<long long>(NI) * NJ * NK;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t % NI) - 1, j = static_cast<int>((t / NI) % NJ) - 1;
    const int k = static_cast<int>(t / (static_cast<long long>(NI) * NJ)) - 1;
    const long long idx = CellIdx(b, i, j, k);
    double r[R::DN];
    LoadCell<neq>(b.state, b.fs, idx, r);
    const MixK<NS> m = MixOf<NS>(p.gas, r);
    const double t0 = r[E::ie] * m.tFac;
    r[R::iH] = m.hf + m.cp * t0 + 0.5 * VelMagSq<NS>(r);  // as MakeIngr
    r[R::iA] = sqrt(m.gamma * r[E::ie] * m.rhoInv);
    r[R::iVt] = 0.0;
    r[R::iVtT] = 0.0;
    if (p.isViscous) {
      // state-dependent factors of the viscous face spectral radii (NeighbourViscTerms)
      const double rho = SpeciesSum<NS>(r);
      const double mu = __ldg(b.viscosity + idx);
      const double mut = NT > 0 ? __ldg(b.eddyVisc + idx) : 0.0;
      r[R::iVt] = ViscSpecFactor(p.tr, rho, Gamma<NS>(p.gas, r), mu, mut);
      if (NT > 0)
        r[R::iVtT] = TurbViscSpecFactor(p.tr.turbModel, p.tr.scaling, rho, r[NS + 4],
                                        r[NS + 4 + (NT > 1 ? 1 : 0)], mu, mut, __ldg(b.f1 + idx));
    }
#pragma unroll
    for (int e = 0; e < neq; ++e) r[R::iB + e] = __ldg(b.rhs + e * b.fs + idx);
    r[R::iD] = __ldg(b.dinv + idx);
    r[R::iD + 1] = NT > 0 ? __ldg(b.dinv + b.fs + idx) : 0.0;
    double2 *o = reinterpret_cast<double2 *>(dyn + t * R::DN);
#pragma unroll
    for (int q = 0; q < R::DN / 2; ++q) o[q] = make_double2(r[2 * q], r[2 * q + 1]);
  }
}

template <int N>
__device__ __forceinline__ void LoadRec(const double *__restrict__ src, double *dst) {
  static_assert(N % 2 == 0, "records are read as 16-byte words");
  const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
  for (int q = 0; q < N / 2; ++q) {
    const double2 v = __ldg(s2 + q);
    dst[2 * q] = v.x;
    dst[2 * q + 1] = v.y;
  }
}

template <int NS, int NT, bool FORWARD, int TJ, int TK>
__global__ void __launch_bounds__(((TJ * TK + 2 * (TJ + TK) + 31) / 32) * 32 + 32, 2)
    LusgsPencilKernel(BlockDev b, Params p, int fullGS, const double *__restrict__ dyn,
                      const double *__restrict__ geo, const int2 *__restrict__ order, int nPencils,
                      int nbJ, WaveSync *sync) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT>;
  constexpr int neq = E::neq, nf = NS + 4;
  constexpr int NCELL = TJ * TK, NH = 2 * (TJ + TK);
  constexpr int NCOMP = ((NCELL + NH + 31) / 32) * 32;  // compute threads; then the flag warp
  constexpr int PJ = TJ + 1, NP = PJ * (TK + 1);        // record positions: jl, kl in [-1, T-1]
  constexpr int NST = neq + 4;                          // s | H a vt vtT
  constexpr int NSTL = (NST + 1) & ~1;                  // ... read as whole 16-byte words
  __shared__ double rec[2][R::RN][NP];
  __shared__ double up[2][neq][NCELL];  // pushes from the j- and k-neighbour ahead
  __shared__ int sTicket;

  const int tid = threadIdx.x;
  const bool isFlag = tid >= NCOMP;
  // role: 0 line of the pencil, 1 / 2 halo behind in j / k, 3 / 4 halo ahead in j / k, 5 idle
  int role = 5, jl = 0, kl = 0;
  if (tid < NCELL) {
    role = 0; jl = tid % TJ; kl = tid / TJ;
  } else if (tid < NCELL + TK) {
    role = 1; jl = -1; kl = tid - NCELL;
  } else if (tid < NCELL + TK + TJ) {
    role = 2; jl = tid - NCELL - TK; kl = -1;
  } else if (tid < NCELL + 2 * TK + TJ) {
    role = 3; kl = tid - NCELL - TK - TJ;
  } else if (tid < NCELL + NH) {
    role = 4; jl = tid - NCELL - 2 * TK - TJ;
  }
  const int nd[3] = {b.ni, b.nj, b.nk};
  int *done = sync->done;
  const long long wRowStride = b.ni + 2;

  for (;;) {
    if (tid == 0) sTicket = static_cast<int>(atomicAdd(&sync->ticket, 1u));
    __syncthreads();
    const int ticket = sTicket;
    if (ticket >= nPencils) return;
    const int2 bc = order[ticket];
    const int J0 = bc.x * TJ, K0 = bc.y * TK;
    const int tj = min(TJ, b.nj - J0), tk = min(TK, b.nk - K0);
    const int nSteps = b.ni + tj + tk - 2;

    if (isFlag) {
      // ---- flag warp: publishes this pencil's progress, waits for the two pencils behind -------
      int *myFlag = done + bc.x + nbJ * bc.y;
      const int *flagJ = bc.x > 0 ? done + (bc.x - 1) + nbJ * bc.y : nullptr;
      const int *flagK = bc.y > 0 ? done + bc.x + nbJ * (bc.y - 1) : nullptr;
      int seenJ = 0, seenK = 0;
      // at local plane q the halo threads read the j-neighbour pencil's plane q + TJ (k: q + TK)
      auto waitFor = [&](int q) {
        if (flagJ)
          while (seenJ < q + TJ + 1) seenJ = LdAcquire(flagJ);
        if (flagK)
          while (seenK < q + TK + 1) seenK = LdAcquire(flagK);
      };
      __syncthreads();
      for (int q = -1; q < nSteps; ++q) {
        if (tid == NCOMP) {
          if (q > 0) StRelease(myFlag, q);  // planes 0 .. q-1 are in global memory
          waitFor(q);                       // ... while the compute threads are in phase A
        }
        __syncthreads();
        __syncthreads();
      }
      if (tid == NCOMP) StRelease(myFlag, kWaveDone);
      continue;
    }

    // ---- compute threads ---------------------------------------------------------------------
    if (role == 3) jl = tj;
    if (role == 4) kl = tk;
    const bool lineValid = role == 0   ? (jl < tj && kl < tk)
                           : role == 1 ? kl < tk
                           : role == 2 ? jl < tj
                           : role == 3 ? kl < tk
                           : role == 4 ? jl < tj
                                       : false;
    const int J = J0 + jl, K = K0 + kl;  // sweep space, -1 .. n
    const int j = FORWARD ? J : b.nj - 1 - J, k = FORWARD ? K : b.nk - 1 - K;
    const long long idxRow = lineValid ? CellIdx(b, 0, j, k) : 0;
    const long long wRow = lineValid ? WaveIdx(b, 0, j, k) : 0;
    const int pos = (jl + 1) + PJ * (kl + 1);
    const int cell = jl + TJ * kl;
    // sweep coordinate I -> cell index offset along the line
    auto iOf = [&](int I) { return FORWARD ? I : b.ni - 1 - I; };

    double pushI[neq];  // contribution of the next cell of this line (ahead in i) to this one
#pragma unroll
    for (int e = 0; e < neq; ++e) pushI[e] = 0.0;

    __syncthreads();
    for (int q = -1; q < nSteps; ++q) {
      const int I = q - jl - kl;
      // a record is a 128-byte line of its own: pull the lines this thread reads kPF cells from
      // now into L2 (updates: one 32-byte sector holds four cells)
      if (p.prefetch && lineValid && role == 0) {
        constexpr int kPF = 8;
        const int X = I + kPF;
        if (X >= 0 && X < b.ni) {
          const int ix = iOf(X);
#pragma unroll
          for (int l = 0; l < (R::DN * 8 + 127) / 128; ++l)
            PrefetchL2(dyn + (wRow + ix) * R::DN + 16 * l);
          PrefetchL2(geo + (wRow + ix) * R::GN);
          if ((X & 3) == 0) {
#pragma unroll
            for (int e = 0; e < neq; ++e) PrefetchL2(b.x + e * b.fs + idxRow + ix);
          }
        }
      }
      // ---------------- phase A: old ingredients of the cell solved next plane, pushes ----------
      if (fullGS && lineValid && (role == 0 || role >= 3)) {
        const int X = I + 1;
        const bool wantJK = X >= 0 && X <= b.ni - 1;
        const bool wantI = role == 0 && X >= 1 && X <= b.ni;
        if (wantJK || wantI) {
          const int ix = iOf(X);
          double st[NSTL], g[R::GN], du[neq], sn[neq], Hn;
          LoadRec<NSTL>(dyn + (wRow + ix) * R::DN, st);
          LoadRec<R::GN>(geo + (wRow + ix) * R::GN, g);
#pragma unroll
          for (int e = 0; e < neq; ++e) du[e] = b.x[e * b.fs + idxRow + ix];
          MakeIngrDyn<NS, NT>(p.gas, st, du, sn, &Hn);
          auto ld = [&](int c) {
            return c < neq + 2 ? st[c]
                               : (c < 2 * neq + 2 ? du[c - neq - 2]
                                                  : (c < 3 * neq + 2 ? sn[c - 2 * neq - 2] : Hn));
          };
          // X is the geometrically upper neighbour of the cells it pushes to in a forward sweep
          if (wantI) {
#pragma unroll
            for (int e = 0; e < neq; ++e) pushI[e] = 0.0;
            OffDiagFromIngr<NS, NT>(ld, g, !FORWARD, pushI, g[12] * st[R::iVt], g[12] * st[R::iVtT]);
          }
          if (wantJK) {
            if (role == 3 || (role == 0 && jl > 0)) {
              double acc[neq];
#pragma unroll
              for (int e = 0; e < neq; ++e) acc[e] = 0.0;
              OffDiagFromIngr<NS, NT>(ld, g + 4, !FORWARD, acc, g[13] * st[R::iVt],
                                      g[13] * st[R::iVtT]);
#pragma unroll
              for (int e = 0; e < neq; ++e) up[0][e][cell - 1] = acc[e];
            }
            if (role == 4 || (role == 0 && kl > 0)) {
              double acc[neq];
#pragma unroll
              for (int e = 0; e < neq; ++e) acc[e] = 0.0;
              OffDiagFromIngr<NS, NT>(ld, g + 8, !FORWARD, acc, g[14] * st[R::iVt],
                                      g[14] * st[R::iVtT]);
#pragma unroll
              for (int e = 0; e < neq; ++e) up[1][e][cell - TJ] = acc[e];
            }
          }
        }
      }
      __syncthreads();
      // ---------------- phase B: solve, new ingredients, record ---------------------------------
      if (lineValid && role <= 2) {
        const bool solve = role == 0 && I >= 0 && I < b.ni;
        // ghost cell behind the line's first cell / cells of the pencils behind: record only
        const bool foreign = (role == 0 && I == -1) || (role != 0 && I >= 0 && I < b.ni);
        if (solve || foreign) {
          const int ic = iOf(I);
          double d[R::DN], xn[neq];
          LoadRec<R::DN>(dyn + (wRow + ic) * R::DN, d);
          if (solve) {
            double g[R::GN];
            LoadRec<R::GN>(geo + (wRow + ic) * R::GN, g);
            const int c[3] = {ic, j, k};
            double bs[neq], as[neq];
#pragma unroll
            for (int e = 0; e < neq; ++e) {
              bs[e] = 0.0;
              as[e] = 0.0;
            }
            // a neighbour contributes if it is a physical cell or lies across a connection
            // (ref src/procBlock.cpp:1064,1115); behind = lower side in a forward sweep
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
              const int d1 = (dd + 1) % 3, d2 = (dd + 2) % 3;
              const bool lo = c[dd] > 0 || ConnAcross(b, 2 * dd + 1, c[d1], nd[d1], c[d2]);
              const bool hi = c[dd] < nd[dd] - 1 || ConnAcross(b, 2 * dd + 2, c[d1], nd[d1], c[d2]);
              const bool useBehind = FORWARD ? lo : hi, useAhead = FORWARD ? hi : lo;
              if (useBehind) {
                const int np = pos - (dd == 0 ? 0 : (dd == 1 ? 1 : PJ));
                const double(*rr)[NP] = rec[(q + 1) & 1];
                auto ld = [&](int cc) { return rr[cc < neq + 2 ? cc : cc + 2][np]; };
                OffDiagFromIngr<NS, NT>(ld, g + 4 * dd, FORWARD, bs, g[12 + dd] * rr[R::iVt][np],
                                        g[12 + dd] * rr[R::iVtT][np]);
              }
              if (useAhead && fullGS) {
                if (dd == 0) {
#pragma unroll
                  for (int e = 0; e < neq; ++e) as[e] += pushI[e];
                } else {
#pragma unroll
                  for (int e = 0; e < neq; ++e) as[e] += up[dd - 1][e][cell];
                }
              }
            }
            // forward: x = D^-1 (b + (L - U)); backward: D^-1 ((b + L) - U), or on the first sweep
            // without initialisation x - D^-1 U (ref src/linearSolver.cpp:341-428)
#pragma unroll
            for (int e = 0; e < neq; ++e) {
              const double dinv = d[R::iD + (e < nf ? 0 : 1)];
              double r;
              if (FORWARD) r = d[R::iB + e] + (bs[e] - as[e]);
              else if (fullGS) r = (d[R::iB + e] + as[e]) - bs[e];
              else r = bs[e];
              r *= dinv;
              if (!FORWARD && !fullGS) r = b.x[e * b.fs + idxRow + ic] - r;
              xn[e] = r;
              __stcg(b.x + e * b.fs + idxRow + ic, r);
            }
          } else {
            // updates of other pencils are rewritten during the sweep: L2, never L1
#pragma unroll
            for (int e = 0; e < neq; ++e) xn[e] = __ldcg(b.x + e * b.fs + idxRow + ic);
          }
          double sn[neq], Hn;
          MakeIngrDyn<NS, NT>(p.gas, d, xn, sn, &Hn);
          double(*rw)[NP] = rec[q & 1];
#pragma unroll
          for (int e = 0; e < NST; ++e) rw[e][pos] = d[e];
#pragma unroll
          for (int e = 0; e < neq; ++e) {
            rw[NST + e][pos] = xn[e];
            rw[NST + neq + e][pos] = sn[e];
          }
          rw[NST + 2 * neq][pos] = Hn;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace aither
