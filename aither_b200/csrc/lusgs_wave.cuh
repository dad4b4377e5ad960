// lusgs_wave.cuh -- LU-SGS / BLU-SGS as ONE persistent wavefront launch per half sweep.
//
// ref: src/linearSolver.cpp:341-428 (lusgs::LUSGS_Forward / LUSGS_Backward), hyperplane order
// src/utility.cpp:377-398, off-diagonals src/procBlock.cpp:1056-1170.
//
// The reference visits the cells in i+j+k hyperplane order; a cell needs the NEW update of its
// three neighbours "behind" it and the OLD update of the three "ahead". Any order that honours
// that gives the same numbers. One launch per hyperplane (LusgsPlaneSplitKernel) pays a grid-wide
// dependent launch per plane: 8 us x (ni+nj+nk) per half sweep, 4 % of the HBM roofline. Here:
//
//   * the block is cut into PENCILS: TJ x TK cells in (j, k), the whole block long in i;
//   * a thread block walks its pencil in LOCAL hyperplanes q = I + jl + kl (one __syncthreads per
//     plane, new updates handed on through shared memory);
//   * pencils depend on the pencil below in j and in k only. Every pencil publishes how many local
//     planes it has finished (st.global + __threadfence by a dedicated warp); its two successors
//     poll that counter and run TJ (TK) + 2 planes behind. Pencils are handed out through an
//     atomic ticket in anti-diagonal order, so a pencil's predecessors are always running or done:
//     no deadlock whatever the number of resident thread blocks.
//   * a cell gets EIGHT lanes as in LusgsPlaneSplitKernel: lanes 0..2 the neighbours behind in
//     i, j, k (they wait for the new update), lanes 3..5 the ones ahead; the six products are
//     summed with shuffles in the reference's order (L: i, j, k; U: i, j, k), then lane e finishes
//     equation e (row e of D^-1 for block matrices).
//   * everything that does not depend on the incoming update is taken OFF the dependent chain: the
//     neighbour's state / face area / viscous data are loaded one plane ahead, and the
//     update-independent half of the off-diagonal (conserved state, old flux, spectral radius; the
//     whole flux Jacobian for block matrices) is formed one plane ahead too ("Prepare"). What is
//     left per plane: U + dU -> primitives -> flux (scalar), or one matrix-vector product (block).
//
// "Sweep space": a backward sweep is the forward sweep of the mirrored block (I = ni-1-i, ...), so
// the kernel is written once; FORWARD only decides how sweep coordinates map to cells and which
// geometric side (lower / upper) is "behind".
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace aither {

struct WaveSync {           // per block, device memory, zeroed before every half sweep
  unsigned int ticket;
  int pad[31];
  int done[1];              // [nPencils] local planes finished by each pencil (sweep-space index)
};

constexpr int kWaveDone = 1 << 30;

// Progress counters are polled with a RELAXED gpu-scope load. ld.acquire.gpu compiles to
// LDG.STRONG + CCTL.IVALL -- every poll throws away the SM's whole L1, which the grid-line
// walkers live on (ncu, profiles/r02c: 12 % of all stall samples on that one instruction, L1 hit
// rate 34 %). Nothing here needs the invalidation: whatever another thread block writes during a
// sweep (the update) is read with ld.cg, i.e. at L2, after the __syncthreads that follows the
// poll; the producer orders its st.cg before the counter with st.release.
__device__ __forceinline__ int LdAcquire(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void StRelease(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one neighbour of one cell: what is loaded (Load), what can be formed without the neighbour's
// update (Prepare), and the product with the update (Product).
template <int NS, int NT, int JAC>
struct WaveNb {
  using E = Eq<NS, NT>;
  static constexpr int neq = E::neq;
  static constexpr bool kBlock = JAC == kJacBlock, kRoe = JAC == kJacRoe;
  // loaded
  double sn[neq], fa[4], dun[neq];
  double dist, mu, mut, f1;
  double vg[kBlock ? 9 : 1];
  double own[kRoe ? neq : 1];
  // prepared. scalar: cons[neq] | fo[neq] | sr | srT      block: J[Blk::n]      roe: oldFlux[neq]
  static constexpr int nA = kBlock ? Blk<NS, NT>::n : (kRoe ? neq : 2 * neq + 2);
  double a[nA];
  bool valid, fromSmem;

  __device__ __forceinline__ void Load(const BlockDev &b, const Params &p, int d, long long idx,
                                       long long nidx, long long fidx, bool needX) {
    LoadCell<neq>(b.state, b.fs, nidx, sn);
#pragma unroll
    for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[d] + q * b.fs + fidx);
    if (needX) {
      // updates are rewritten by other thread blocks during the sweep: L2, never L1
#pragma unroll
      for (int e = 0; e < neq; ++e) dun[e] = __ldcg(b.x + e * b.fs + nidx);
    }
    if (p.isViscous) {
      dist = __ldg(b.dist[d] + fidx);
      mu = __ldg(b.viscosity + nidx);
      mut = NT > 0 ? __ldg(b.eddyVisc + nidx) : 0.0;
      f1 = NT > 0 ? __ldg(b.f1 + nidx) : 0.0;
      if (kBlock) {
#pragma unroll
        for (int q = 0; q < 9; ++q) vg[q] = __ldg(b.velGrad + q * b.fs + nidx);
      }
    }
    if (kRoe) LoadCell<neq>(b.state, b.fs, idx, own);
  }

  // ref: src/fluxJacobian.cpp:122-194,240-296; the same expressions as OffDiagOne (kernels.cuh)
  __device__ __forceinline__ void Prepare(const Params &p, bool positive) {
    if constexpr (kBlock) {
      RusanovFluxJacobian<NS, NT>(p.gas, sn, fa, positive, a);
      if (p.isViscous) {
        double V[Blk<NS, NT>::n];
        ApproxTslJacobian<NS, NT>(p.gas, p.tr, sn, mu, mut, f1, fa, dist, positive, vg, V);
#pragma unroll
        for (int q = 0; q < Blk<NS, NT>::n; ++q) a[q] = positive ? a[q] - V[q] : a[q] + V[q];
      }
    } else if constexpr (kRoe) {
      RoeFlux<NS, NT>(p.gas, sn, own, fa, a);
    } else {
      PrimToCons<NS, NT>(p.gas, sn, a);
      PhysicalFlux<NS, NT>(p.gas, sn, fa, a + neq);
      double extra = 0.0, extraT = 0.0;
      if (p.isViscous) {
        const double length = fa[3] / dist;
        const double rho = SpeciesSum<NS>(sn);
        extra = length * ViscSpecFactor(p.tr, rho, Gamma<NS>(p.gas, sn), mu, mut);
        if (NT > 0)
          extraT = length * TurbViscSpecFactor(p.tr.turbModel, p.tr.scaling, rho, sn[NS + 4],
                                               sn[NS + 4 + (NT > 1 ? 1 : 0)], mu, mut, f1);
      }
      a[2 * neq] = InvFaceSpectralRadius<NS>(sn, SoS<NS>(p.gas, sn), fa) + extra;
      double srT = 0.0;
      if (NT > 0) {
        const double velNorm = sn[NS] * fa[0] + sn[NS + 1] * fa[1] + sn[NS + 2] * fa[2];
        srT = (positive ? 0.5 * fa[3] * fabs(velNorm + fabs(velNorm))
                        : 0.5 * fa[3] * fabs(velNorm - fabs(velNorm))) + extraT;
      }
      a[2 * neq + 1] = srT;
    }
  }

  __device__ __forceinline__ void Product(const Params &p, const double *du, bool positive,
                                          double *od) const {
    if constexpr (kBlock) {
      BlockMult<NS, NT>(a, du, od);
    } else if constexpr (kRoe) {
      double newFlux[neq], su[neq];
      UpdatePrimWithCons<NS, NT>(p.gas, sn, du, su);
      if (positive) RoeFlux<NS, NT>(p.gas, su, own, fa, newFlux);
      else RoeFlux<NS, NT>(p.gas, own, su, fa, newFlux);
#pragma unroll
      for (int e = 0; e < neq; ++e) od[e] = fa[3] * (newFlux[e] - a[e]);
    } else {
      double su[neq], fn[neq];
      UpdatePrimFromCons<NS, NT>(p.gas, a, du, su);
      PhysicalFlux<NS, NT>(p.gas, su, fa, fn);
      const double sr = a[2 * neq], srT = a[2 * neq + 1];
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        const double fc = e < NS + 4 ? 0.5 * fa[3] * (fn[e] - a[neq + e]) : 0.0;
        const double srd = (e < NS + 4 ? sr : srT) * du[e];
        od[e] = positive ? fc + srd : fc - srd;
      }
    }
  }
};

template <int NS, int NT, bool FORWARD, int JAC, int TJ, int TK>
__global__ void __launch_bounds__(TJ *TK * 8 + 32, 1)
    LusgsWaveKernel(BlockDev b, Params p, int fullGS, const int2 *__restrict__ order, int nPencils,
                    int nbJ, WaveSync *sync) {
  using E = Eq<NS, NT>;
  using Nb = WaveNb<NS, NT, JAC>;
  constexpr int neq = E::neq, NC = TJ * TK, NTHR = NC * 8;
  constexpr int nf = NS + 4;
  constexpr bool kBlock = JAC == kJacBlock;
  static_assert(!kBlock || neq <= 8, "block rows are finished by the eight lanes of a cell");
  __shared__ double sx[2][neq][NC];
  __shared__ int sTicket;

  const int tid = threadIdx.x;
  const bool isSync = tid >= NTHR;
  const int lane8 = tid & 7;
  const int cell = isSync ? 0 : tid >> 3;
  const int jl = cell % TJ, kl = cell / TJ;
  const int d = lane8 % 3;
  const bool behind = lane8 < 3;
  const bool hasTask = lane8 < 6 && (behind || fullGS != 0);
  const bool lowerSide = behind == FORWARD;  // geometric side of this lane's neighbour
  const long long st = Stride(b, d);
  const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
  const int nd[3] = {b.ni, b.nj, b.nk};
  const int srcCell = cell - (d == 0 ? 0 : (d == 1 ? 1 : TJ));
  const int base = (tid & 31) & ~7;
  int *done = sync->done;

  for (;;) {
    if (tid == 0) sTicket = static_cast<int>(atomicAdd(&sync->ticket, 1u));
    __syncthreads();
    const int ticket = sTicket;
    if (ticket >= nPencils) return;
    const int2 bc = order[ticket];
    const int J0 = bc.x * TJ, K0 = bc.y * TK;
    const int tj = min(TJ, b.nj - J0), tk = min(TK, b.nk - K0);
    const int nSteps = b.ni + tj + tk - 2;

    if (isSync) {
      // ---- flag warp: publishes this pencil's progress, waits for the two pencils behind -------
      int *myFlag = done + bc.x + nbJ * bc.y;
      const int *flagJ = bc.x > 0 ? done + (bc.x - 1) + nbJ * bc.y : nullptr;
      const int *flagK = bc.y > 0 ? done + bc.x + nbJ * (bc.y - 1) : nullptr;
      int seenJ = 0, seenK = 0;
      // our local plane q reads the j-neighbour pencil's local plane q + TJ - 1 (k: q + TK - 1)
      auto waitFor = [&](int q) {
        if (flagJ)
          while (seenJ < q + TJ) seenJ = LdAcquire(flagJ);
        if (flagK)
          while (seenK < q + TK) seenK = LdAcquire(flagK);
      };
      if (tid == NTHR) waitFor(1);
      __syncthreads();
      for (int q = 0; q < nSteps; ++q) {
        if (tid == NTHR) {
          if (q > 0) StRelease(myFlag, q);  // planes 0 .. q-1 are in global memory
          waitFor(q + 2);
        }
        __syncthreads();
      }
      if (tid == NTHR) StRelease(myFlag, kWaveDone);
      continue;
    }

    // ---- compute warps ---------------------------------------------------------------------
    const int J = J0 + jl, K = K0 + kl;
    const bool cellValid = jl < tj && kl < tk;
    const int j = FORWARD ? J : b.nj - 1 - J, k = FORWARD ? K : b.nk - 1 - K;
    const long long idxRow = CellIdx(b, 0, cellValid ? j : 0, cellValid ? k : 0);

    auto loadNb = [&](int q, Nb &r) {
      const int I = q - jl - kl;
      r.valid = false;
      r.fromSmem = false;
      if (!(hasTask && cellValid && I >= 0 && I < b.ni)) return;
      const int i = FORWARD ? I : b.ni - 1 - I;
      const int c[3] = {i, j, k};
      // a neighbour contributes if it is a physical cell or lies across a connection
      // (ref src/procBlock.cpp:1064,1115)
      const bool contributes =
          lowerSide ? (c[d] > 0 || ConnAcross(b, 2 * d + 1, c[d1], nd[d1], c[d2]))
                    : (c[d] < nd[d] - 1 || ConnAcross(b, 2 * d + 2, c[d1], nd[d1], c[d2]));
      if (!contributes) return;
      r.valid = true;
      const long long idx = idxRow + i;
      // behind and inside the pencil: the update arrives through shared memory one plane later
      r.fromSmem = behind && (d == 0 ? I > 0 : (d == 1 ? jl > 0 : kl > 0));
      r.Load(b, p, d, idx, lowerSide ? idx - st : idx + st, lowerSide ? idx : idx + st,
             !r.fromSmem);
    };

    __syncthreads();  // the flag warp has seen the planes our first two steps read
    Nb cur, nxt;
    loadNb(0, cur);
    if (cur.valid) cur.Prepare(p, lowerSide);

    for (int q = 0; q < nSteps; ++q) {
      const int I = q - jl - kl;
      const bool active = cellValid && I >= 0 && I < b.ni;
      const int i = FORWARD ? I : b.ni - 1 - I;
      const long long idx = idxRow + (active ? i : 0);
      // pull the lines the next planes read into L2, one request per 32-byte sector
      if (p.prefetch && active && (I & 3) == 0 && I + 16 < b.ni) {
        const long long ahead = FORWARD ? 16 : -16;
        if (lane8 < 6) {
          const long long nidx = (lowerSide ? idx - st : idx + st) + ahead;
          const long long fidx = (lowerSide ? idx : idx + st) + ahead;
#pragma unroll
          for (int e = 0; e < neq; ++e) PrefetchL2(b.state + e * b.fs + nidx);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) PrefetchL2(b.fA[d] + qq * b.fs + fidx);
          if (p.isViscous) {
            PrefetchL2(b.dist[d] + fidx);
            PrefetchL2(b.viscosity + nidx);
          }
        } else if (lane8 == 6) {
#pragma unroll
          for (int e = 0; e < neq; ++e) PrefetchL2(b.rhs + e * b.fs + idx + ahead);
        } else {
          if (kBlock) {
#pragma unroll
            for (int qq = 0; qq < Blk<NS, NT>::n; ++qq) PrefetchL2(b.dinv + qq * b.fs + idx + ahead);
          } else {
            PrefetchL2(b.dinv + idx + ahead);
            if (NT > 0) PrefetchL2(b.dinv + b.fs + idx + ahead);
          }
        }
      }
      // next plane's neighbour data: in flight while this plane is finished
      loadNb(q + 1, nxt);
      // this lane's equation(s) of the cell: right-hand side, diagonal (row), old update
      double bOwn[2] = {0.0, 0.0}, xOld[2] = {0.0, 0.0}, dOwn[kBlock ? nf : 2];
      if (active) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = lane8 + 8 * h;
          if (e < neq) {
            bOwn[h] = __ldg(b.rhs + e * b.fs + idx);
            if (!FORWARD && !fullGS) xOld[h] = __ldcg(b.x + e * b.fs + idx);
            if (!kBlock) dOwn[h] = __ldg(b.dinv + (e < nf ? 0 : 1) * b.fs + idx);
          }
        }
        if (kBlock) {
          if (lane8 < nf) {
#pragma unroll
            for (int cc = 0; cc < nf; ++cc) dOwn[cc] = __ldg(b.dinv + (lane8 * nf + cc) * b.fs + idx);
          } else if (lane8 < neq) {
#pragma unroll
            for (int cc = 0; cc < NT; ++cc)
              dOwn[cc] = __ldg(b.dinv + (nf * nf + (lane8 - nf) * NT + cc) * b.fs + idx);
          }
        }
      }

      // ---- the dependent chain: neighbour's update -> product -> sums -> new update ----------
      double od[neq];
#pragma unroll
      for (int e = 0; e < neq; ++e) od[e] = 0.0;
      if (cur.valid) {
        double du[neq];
        if (cur.fromSmem) {
#pragma unroll
          for (int e = 0; e < neq; ++e) du[e] = sx[(q + 1) & 1][e][srcCell];
        } else {
#pragma unroll
          for (int e = 0; e < neq; ++e) du[e] = cur.dun[e];
        }
        cur.Product(p, du, lowerSide, od);
      }
      // L = ((0 + od_i) + od_j) + od_k from the lower-side lanes, U from the upper-side ones
      // (ref src/procBlock.cpp:1056-1170); in a backward sweep lanes 0..2 hold the upper side
      double own[2] = {0.0, 0.0};
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
          s0 += __shfl_sync(0xffffffffu, od[e], base + qq);
          s1 += __shfl_sync(0xffffffffu, od[e], base + 3 + qq);
        }
        const double L = FORWARD ? s0 : s1, U = FORWARD ? s1 : s0;
        double r;
        if (FORWARD) r = bOwn[e >> 3] + (L - U);
        else if (fullGS) r = (bOwn[e >> 3] + L) - U;
        else r = U;
        if ((e & 7) == lane8) own[e >> 3] = r;
      }
      if (kBlock) {
        // row lane8 of D^-1 times the vector the eight lanes hold (DiagMult's order)
        double acc = 0.0;
        if (lane8 < nf) {
#pragma unroll
          for (int cc = 0; cc < nf; ++cc)
            acc += dOwn[cc] * __shfl_sync(0xffffffffu, own[0], base + cc);
        } else {
#pragma unroll
          for (int cc = 0; cc < nf; ++cc) (void)__shfl_sync(0xffffffffu, own[0], base + cc);
        }
        double accT = 0.0;
#pragma unroll
        for (int cc = 0; cc < NT; ++cc) {
          const double v = __shfl_sync(0xffffffffu, own[0], base + nf + cc);
          if (lane8 >= nf && lane8 < neq) accT += dOwn[cc] * v;
        }
        own[0] = lane8 < nf ? acc : accT;
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) own[h] = own[h] * dOwn[h];
      }
      if (active) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = lane8 + 8 * h;
          if (e < neq) {
            const double xn = (!FORWARD && !fullGS) ? xOld[h] - own[h] : own[h];
            sx[q & 1][e][cell] = xn;
            __stcg(b.x + e * b.fs + idx, xn);
          }
        }
      }
      // ---- off the chain: next plane's update-independent half --------------------------------
      if (nxt.valid) nxt.Prepare(p, lowerSide);
      cur = nxt;
      __syncthreads();
    }
  }
}

}  // namespace aither
