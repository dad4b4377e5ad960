// implicit_tma.cuh -- the implicit sweep (DPLUR / matrix residual) as a TMA-fed plane march.
//
// Same mathematics and accumulation order as ImplicitMarchKernel (march.cuh); what changes is how
// the bytes reach the SM. ncu on the register-fed version (profiles/r01c_*): 34 % of DRAM peak,
// 24 % occupancy, "long scoreboard" the dominant stall -- the loads of a plane were issued next to
// their uses and only ~5 were in flight per warp. Here one elected thread issues, per plane, four
// bulk tensor copies (state and update tiles with a one-cell halo, i- and j-face areas) into a
// double-buffered shared-memory stage, ONE PLANE AHEAD of the plane being computed; the copies
// complete on an mbarrier, hold no registers, and keep ~84 KB per SM in flight. What has no
// in-plane reuse (k-face areas, right-hand side, diagonal) is loaded to registers one plane ahead.
//
//   block   : 32 x 16 cells, 512 threads, 1 block per SM (207 KB of shared memory)
//   stage s : state [5][18][36] | update [5][18][36] | i-face areas [4][16][34] | j-face [4][17][32]
//   sG      : per-cell ingredients the neighbours need beyond state and update:
//             H, a, updated primitive state (5), updated H  -- computed once per cell per sweep
#pragma once
#include <cuda_runtime.h>

#include "march.cuh"
#include "tma.cuh"

namespace aither {

constexpr int kQI = 32, kQJ = 16, kQThreads = kQI * kQJ;
// state / update tile with a one-cell halo. TMA needs the first element of a box 16-byte aligned
// in global memory (an odd double coordinate traps as an illegal instruction), so the tile starts
// TWO cells left of the block's first column and is 36 wide; column 0 and 35 are never read.
constexpr int kQL = 2;
constexpr int kQPI = kQI + 2 * kQL, kQPJ = kQJ + 2, kQPC = kQPI * kQPJ;

constexpr int kQAI = kQI + 2;  // i-face tile pitch: 33 faces, padded to a 16-byte multiple
template <int NS, int NT>
struct ImplTma {
  static constexpr int neq = NS + 4 + NT;
  static constexpr int nG = neq + 3;                       // H, a, sn[neq], Hn
  static constexpr int szS = ((neq * kQPC + 15) / 16) * 16;  // doubles, 128-byte multiple
  static constexpr int szAi = 4 * kQJ * kQAI;              // [4][16][34]
  static constexpr int szAj = 4 * (kQJ + 1) * kQI;         // [4][17][32]
  static_assert(szAi % 16 == 0 && szAj % 16 == 0, "TMA destinations must stay 128-byte aligned");
  static constexpr int stage = 2 * szS + szAi + szAj;
  static constexpr int szG = nG * kQPC;
  static constexpr size_t bytes = sizeof(double) * (2 * stage + szG) + 64;
  static constexpr unsigned txState = sizeof(double) * 2 * neq * kQPC;
  static constexpr unsigned txFaces = sizeof(double) * (szAi + szAj);
};

struct ImplMaps {
  CUtensorMap cell;   // box (36, 18, 1, neq)
  CUtensorMap faceI;  // box (34, 16, 1, 4)
  CUtensorMap faceJ;  // box (32, 17, 1, 4)
};

// MODE kModeDplur: xout = D^-1 (b + L(xin) - U(xin))      (ref src/linearSolver.cpp:473-507)
// MODE kModeAxmb : mr = -((D x - (L - U)) - b), partial sums of mr^2 per block (:58-109)
// fState / fX / fAi / fAj: field indices (pointer offset / fs) inside the block's allocation
template <int NS, int NT, int MODE>
__global__ void __launch_bounds__(kQThreads, 1)
    ImplicitTmaKernel(const __grid_constant__ ImplMaps maps, BlockDev b, Params p,
                      const double *__restrict__ xin, double *__restrict__ xout, int fX, int fAi,
                      int fAj, int kChunk, double *__restrict__ partials, int storeField,
                      const int *__restrict__ tiles = nullptr, int tilesX = 0, int tilesY = 0,
                      int nSignal = 0, unsigned int *__restrict__ signal = nullptr, int fState = 0,
                      double *__restrict__ stateOut = nullptr) {
  using E = Eq<NS, NT>;
  using T = ImplTma<NS, NT>;
  constexpr int neq = E::neq;
  extern __shared__ __align__(128) double smem[];
  double *sG = smem + 2 * T::stage;
  uint64_t *full = reinterpret_cast<uint64_t *>(sG + T::szG);

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + kQI * ty;
  // which (column, k-chunk) this thread block works on: its own grid position, or an entry of a
  // tile list (linear tile id = x + X (y + Y z)). The list puts the tiles next to a connected
  // block face first; each of those first `nSignal` thread blocks bumps `signal` when its part of
  // the new update is in memory, and the ghost exchange -- waiting on that counter on another
  // stream -- runs while the remaining tiles are computed.
  int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, gx = gridDim.x, gy = gridDim.y;
  if (tiles != nullptr) {
    const int t = __ldg(tiles + blockIdx.x);
    gx = tilesX;
    gy = tilesY;
    bx = t % gx;
    by = (t / gx) % gy;
    bz = t / (gx * gy);
  }
  const int i0 = bx * kQI, j0 = by * kQJ;
  const int k0 = bz * kChunk;
  const int k1 = min(k0 + kChunk, b.nk);
  const int i = i0 + tx, j = j0 + ty;
  const bool colValid = i < b.ni && j < b.nj;
  const int pc = (tx + kQL) + kQPI * (ty + 1);

  // halo cell this thread also prepares (threads 0..95 cover the ring around the tile)
  int hc = -1;
  if (tid < 2 * kQI) {
    hc = kQL + (tid % kQI) + kQPI * (tid < kQI ? 0 : kQPJ - 1);
  } else if (tid < 2 * kQI + 2 * kQJ) {
    const int q = tid - 2 * kQI;
    hc = (q < kQJ ? kQL - 1 : kQL + kQI) + kQPI * (1 + (q % kQJ));
  }

  const int nIter = k1 - k0 + 2;  // planes k0-1 .. k1
  auto issue = [&](int it) {
    // plane k0 - 1 + it into stage it & 1; the two end planes only feed k-neighbours
    const int k = k0 - 1 + it;
    double *st = smem + static_cast<size_t>(it & 1) * T::stage;
    uint64_t *bar = full + (it & 1);
    const bool interior = it > 0 && it < nIter - 1;
    MbarExpectTx(bar, T::txState + (interior ? T::txFaces : 0u));
    TmaLoad4D(st, &maps.cell, i0 - kQL + b.lp, j0 - 1 + b.g, k + b.g, fState, bar);
    TmaLoad4D(st + T::szS, &maps.cell, i0 - kQL + b.lp, j0 - 1 + b.g, k + b.g, fX, bar);
    if (interior) {
      TmaLoad4D(st + 2 * T::szS, &maps.faceI, i0 + b.lp, j0 + b.g, k + b.g, fAi, bar);
      TmaLoad4D(st + 2 * T::szS + T::szAi, &maps.faceJ, i0 + b.lp, j0 + b.g, k + b.g, fAj, bar);
    }
  };
  if (tid == 0) {
    MbarInit(full, 1);
    MbarInit(full + 1, 1);
    MbarInitFence();
    issue(0);
    issue(1);
  }
  __syncthreads();

  double accLp[neq], accUp[neq];  // pending cell (k-1): complete L, U without the k+1 term
  double carryL[neq];             // L-term for this plane's cell, produced one plane below
  double duPrev[neq];             // update of the pending cell (matrix residual only)
  double sq = 0.0;
#pragma unroll
  for (int e = 0; e < neq; ++e) {
    accLp[e] = 0.0;
    accUp[e] = 0.0;
    carryL[e] = 0.0;
    duPrev[e] = 0.0;
  }
  // what has no in-plane reuse comes through registers: the k-face area above this plane's cell
  // (kept one plane, it is the next plane's lower face), right-hand side and D^-1 / D of the
  // pending cell; the loads are issued at the top of a plane, before the wait on its tiles
  const long long idxCol = CellIdx(b, min(i, b.ni - 1), min(j, b.nj - 1), 0);
  double faCur[4] = {0.0, 0.0, 0.0, 0.0};

  for (int it = 0; it < nIter; ++it) {
    const int k = k0 - 1 + it;
    const double *st = smem + static_cast<size_t>(it & 1) * T::stage;
    const double *sS = st, *sX = st + T::szS, *sAi = st + 2 * T::szS,
                 *sAj = st + 2 * T::szS + T::szAi;
    const bool planeInterior = it > 0 && it < nIter - 1;
    const long long idx = idxCol + static_cast<long long>(k) * b.sk;

    // next plane's register operands: into L2 one plane ahead (one request per 128-byte line)
    if (p.prefetch && (threadIdx.x & 15) == 0 && k + p.prefetch <= b.nk) {
      const long long ahead = static_cast<long long>(p.prefetch) * b.sk;
#pragma unroll
      for (int q = 0; q < 4; ++q) PrefetchL2(b.fA[2] + q * b.fs + idx + b.sk + ahead);
#pragma unroll
      for (int e = 0; e < neq; ++e) PrefetchL2(b.rhs + e * b.fs + idx - b.sk + ahead);
      PrefetchL2((MODE == kModeDplur ? b.dinv : b.diag) + idx - b.sk + ahead);
    }
    // this plane's register operands, issued before waiting on its tiles
    double faUp[4], rhsN[neq], dN = 0.0;
    if (it < nIter - 1) {  // face k + 1: upper face of this plane's cell
#pragma unroll
      for (int q = 0; q < 4; ++q) faUp[q] = __ldg(b.fA[2] + q * b.fs + idx + b.sk);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) faUp[q] = 0.0;
    }
    if (it >= 2) {  // the pending cell (k - 1) is finished in this plane
#pragma unroll
      for (int e = 0; e < neq; ++e) rhsN[e] = __ldg(b.rhs + e * b.fs + idx - b.sk);
      dN = __ldg((MODE == kModeDplur ? b.dinv : b.diag) + idx - b.sk);
    } else {
#pragma unroll
      for (int e = 0; e < neq; ++e) rhsN[e] = 0.0;
    }

    MbarWait(full + (it & 1), (it >> 1) & 1);

    // ---- ingredients of this thread's cell in plane k -------------------------------------
    double s[neq], du[neq], H, a, sn[neq], Hn;
#pragma unroll
    for (int e = 0; e < neq; ++e) {
      s[e] = sS[e * kQPC + pc];
      du[e] = sX[e * kQPC + pc];
    }
    MakeIngr<NS, NT>(p.gas, s, du, &H, &a, sn, &Hn);
    if (MODE == kModeAxmb && stateOut != nullptr && colValid && planeInterior) {
      // the updated primitive state (U + dU -> primitives; ref src/procBlock.cpp:902-915,
      // include/primitive.hpp:206-231) is what MakeIngr has just formed for the off-diagonals:
      // the state is advanced here, into the alternate buffer (the neighbours' old state is
      // still being read), and the update kernel is left with the residual norms
#pragma unroll
      for (int e = 0; e < neq; ++e) stateOut[e * b.fs + idx] = sn[e];
    }
    if (it < nIter - 1) {  // every plane that has a cell above it in the chunk
      sG[pc] = H;
      sG[kQPC + pc] = a;
#pragma unroll
      for (int e = 0; e < neq; ++e) sG[(2 + e) * kQPC + pc] = sn[e];
      sG[(2 + neq) * kQPC + pc] = Hn;
    }
    // own ingredients in the OffDiagFromIngr layout: s | H a | du | sn | Hn
    auto ldOwn = [&](int q) {
      return q < neq ? s[q]
                     : (q == neq ? H
                                 : (q == neq + 1 ? a
                                                 : (q < 2 * neq + 2 ? du[q - neq - 2]
                                                                    : (q < 3 * neq + 2
                                                                           ? sn[q - 2 * neq - 2]
                                                                           : Hn))));
    };
    if (colValid && it >= 2) {
      // U-term of the cell below (k-1) across face k, then finish that cell
      const bool useKhi = k < b.nk || ConnAcross(b, 6, i, b.ni, j);
      if (useKhi) OffDiagFromIngr<NS, NT>(ldOwn, faCur, false, accUp);
      const long long idxm = idx - b.sk;
      if (MODE == kModeDplur) {
#pragma unroll
        for (int e = 0; e < neq; ++e)
          xout[e * b.fs + idxm] = ((rhsN[e] + 0.0) + (accLp[e] - accUp[e])) * dN;
      } else {
#pragma unroll
        for (int e = 0; e < neq; ++e) {
          const double ax = duPrev[e] * dN;
          const double mr = 0.0 - ((ax - (accLp[e] - accUp[e])) - rhsN[e]);
          if (storeField) b.mres[e * b.fs + idxm] = mr;
          sq += mr * mr;
        }
      }
    }
    // ring of halo cells (threads 0..95), after the pending cell has been finished so that this
    // thread's own ingredients no longer occupy registers
    if (planeInterior && hc >= 0) {
      double hs[neq], hdu[neq], hH, ha, hsn[neq], hHn;
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        hs[e] = sS[e * kQPC + hc];
        hdu[e] = sX[e * kQPC + hc];
      }
      MakeIngr<NS, NT>(p.gas, hs, hdu, &hH, &ha, hsn, &hHn);
      sG[hc] = hH;
      sG[kQPC + hc] = ha;
#pragma unroll
      for (int e = 0; e < neq; ++e) sG[(2 + e) * kQPC + hc] = hsn[e];
      sG[(2 + neq) * kQPC + hc] = hHn;
    }
    __syncthreads();  // sG of this plane is complete
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
      auto nb = [&](int c) {
        return [=](int q) {
          return q < neq ? sS[q * kQPC + c]
                         : (q < neq + 2 ? sG[(q - neq) * kQPC + c]
                                        : (q < 2 * neq + 2 ? sX[(q - neq - 2) * kQPC + c]
                                                           : sG[(q - 2 * neq) * kQPC + c]));
        };
      };
      double fa[4];
      const int fi = tx + kQAI * ty, fj = tx + kQI * ty;
      if (useIlo) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = sAi[q * (kQJ * kQAI) + fi];
        OffDiagFromIngr<NS, NT>(nb(pc - 1), fa, true, accLp);
      }
      if (useJlo) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = sAj[q * ((kQJ + 1) * kQI) + fj];
        OffDiagFromIngr<NS, NT>(nb(pc - kQPI), fa, true, accLp);
      }
      if (useKlo) {  // produced from the cell below at the previous plane
#pragma unroll
        for (int e = 0; e < neq; ++e) accLp[e] += carryL[e];
      }
      if (useIhi) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = sAi[q * (kQJ * kQAI) + fi + 1];
        OffDiagFromIngr<NS, NT>(nb(pc + 1), fa, false, accUp);
      }
      if (useJhi) {
#pragma unroll
        for (int q = 0; q < 4; ++q) fa[q] = sAj[q * ((kQJ + 1) * kQI) + fj + kQI];
        OffDiagFromIngr<NS, NT>(nb(pc + kQPI), fa, false, accUp);
      }
    }
    // L-term this cell contributes to the cell above (k+1), across face k+1. Its ingredients are
    // read back from shared memory so that they need not stay in registers through the
    // neighbour phase; the previous carry has been consumed above.
#pragma unroll
    for (int e = 0; e < neq; ++e) carryL[e] = 0.0;
    if (colValid && it + 1 < nIter - 1) {
      auto own = [=](int q) {
        return q < neq ? sS[q * kQPC + pc]
                       : (q < neq + 2 ? sG[(q - neq) * kQPC + pc]
                                      : (q < 2 * neq + 2 ? sX[(q - neq - 2) * kQPC + pc]
                                                         : sG[(q - 2 * neq) * kQPC + pc]));
      };
      OffDiagFromIngr<NS, NT>(own, faUp, true, carryL);
    }
    if (MODE == kModeAxmb) {
#pragma unroll
      for (int e = 0; e < neq; ++e) duPrev[e] = sX[e * kQPC + pc];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) faCur[q] = faUp[q];
    __syncthreads();  // every reader of this stage and of sG is done
    if (tid == 0 && it + 2 < nIter) issue(it + 2);
  }
  if (MODE == kModeAxmb) {
    const int blockLinear = bx + gx * (by + gy * bz);
    BlockSumToPartials<1>(&sq, partials, blockLinear, tid, kQThreads);
  }
  if (signal != nullptr && static_cast<int>(blockIdx.x) < nSignal) {
    __syncthreads();  // every thread's stores of this tile are issued
    if (tid == 0) {
      __threadfence();
      atomicAdd(signal, 1u);
    }
  }
}

}  // namespace aither
