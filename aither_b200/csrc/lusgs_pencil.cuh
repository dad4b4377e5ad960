// lusgs_pencil.cuh -- scalar-diagonal LU-SGS as a persistent pencil wavefront with shared per-cell
// ingredients: one launch per half sweep.
//
// ref: src/linearSolver.cpp:341-428 (lusgs::LUSGS_Forward / LUSGS_Backward), hyperplane order
// src/utility.cpp:377-398, off-diagonals src/procBlock.cpp:1056-1170,
// src/fluxJacobian.cpp:122-162 (RusanovScalarOffDiagonal).
//
// Why the per-hyperplane kernels (and the eight-lanes-per-cell wavefront, lusgs_wave.cuh) sit at
// 4 % of the HBM roofline: every cell evaluates SIX complete off-diagonal products (U + dU ->
// primitives, two fluxes, IEEE divisions: ~4 500 fp64 instructions per cell against ~350 in the
// DPLUR march) -- they are bound by the fp64 pipe, not only by launch latency. What this file does:
//
//   1. the arithmetic of the DPLUR march: per cell the update-dependent ingredients (updated
//      primitive state, its enthalpy: MakeIngrDyn) and one ~40-FMA product per neighbour
//      (OffDiagFromIngr); what does not depend on the update (state, H, a, viscous spectral
//      factors, b, D^-1) is packed once per iteration into a record per cell;
//   2. "sweep space": a backward sweep is the forward sweep of the mirrored block. A cell needs the
//      NEW update of its three neighbours behind it and the OLD update of the three ahead. The
//      ahead-neighbours still hold their pre-sweep update when the cell is solved, so their sum
//      does not depend on the sweep: LusgsAheadKernel forms it for all cells in parallel before
//      the wavefront starts (U of the forward sweep, L of the backward one);
//   3. the wavefront proper (LusgsPencilKernel): the block is cut into pencils of 8 x 7 grid
//      lines; a thread block walks its pencil in local planes q = i + jl + kl, one __syncthreads
//      per plane, the new update handed on through shared memory; pencils are ordered by atomic
//      tickets along anti-diagonals and wait on the progress counters of the two pencils behind
//      them (lusgs_wave.cuh), so a predecessor is always running or done;
//   4. a PLANE-MAJOR workspace: the records of all cells of one plane of one pencil are contiguous
//      in memory, so a plane arrives in shared memory by three bulk copies of the copy engine
//      (cp.async.bulk, a ring of stages several planes ahead) and the walkers read shared memory
//      only. Measured on the way here (lone pencil, us per plane; profiles/r02*): cell-major
//      records read by the walking threads 3.0 -> 1.2 however the work was split over threads --
//      the time tracked the number of distinct 128-byte lines requested per plane (each a
//      separate L1 wavefront), not instructions or bytes.
#pragma once
#include <cuda_runtime.h>

#include "lusgs_wave.cuh"
#include "tma.cuh"

namespace aither {

// pencil cross-section and resident thread blocks per SM (tunable at build time for A/B runs):
// 8 x 8 = 64 cells x 4 lanes + a service warp = 288 threads, one thread block per SM (192^3 x4 sweeps:
// 8x7 18.4 ms, 8x8 17.4, 4x8 and 8x4 with two blocks per SM 18.5 / 19.1)
#ifndef AITHER_PENCIL_TJ
#define AITHER_PENCIL_TJ 8
#endif
#ifndef AITHER_PENCIL_TK
#define AITHER_PENCIL_TK 8
#endif
#ifndef AITHER_PENCIL_CTAS
#define AITHER_PENCIL_CTAS 1
#endif
constexpr int kPTJ = AITHER_PENCIL_TJ, kPTK = AITHER_PENCIL_TK, kPCells = kPTJ * kPTK;
constexpr int kPencilCtasPerSm = AITHER_PENCIL_CTAS;

template <int NS, int NT>
struct PencilRec {
  static constexpr int neq = NS + 4 + NT;
  // per-iteration record: s[neq] | H a vt vtT | b[neq] | dinv dinvT, padded to a stride that is a
  // multiple of 16 bytes and NOT of 128 (records of neighbouring cells in different banks)
  static constexpr int nUsed = 2 * neq + 6;
  static constexpr int nEven = (nUsed + 1) & ~1;
  static constexpr int DN = nEven % 16 == 0 ? nEven + 2 : nEven;
  static constexpr int iH = neq, iA = neq + 1, iVt = neq + 2, iVtT = neq + 3, iB = neq + 4,
                       iD = 2 * neq + 4;
  static constexpr int NST = neq + 4;  // record head: s | H a vt vtT
  // per-block record: behind-side faces i, j, k {nx, ny, nz, |A|} | |A| / dist for i, j, k | pad
  static constexpr int GN = 18;
  static constexpr int AN = (neq + 1) & ~1;  // ahead-sum, 16-byte words
};

// plane-major slot of a cell: pencils tile (j, k) from 0; plane q = i + jl + kl of its pencil
struct PencilLattice {
  int nbJ, nbK, planesPer;  // planesPer = ni + kPTJ + kPTK - 2
};
__host__ __device__ __forceinline__ long long PencilSlot(const PencilLattice &L, int i, int j, int k) {
  const int bJ = j / kPTJ, jl = j - bJ * kPTJ, bK = k / kPTK, kl = k - bK * kPTK;
  return (static_cast<long long>(bJ + L.nbJ * bK) * L.planesPer + (i + jl + kl)) * kPCells +
         (jl + kPTJ * kl);
}

// record head of a cell from the block's fields: what WaveDynKernel packs for the cells of the
// block, formed on the fly (same expressions) for ghost cells and cells of other pencils
template <int NS, int NT>
__device__ __forceinline__ void MakeHead(const BlockDev &b, const Params &p, long long idx,
                                         double *r) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT>;
  LoadCell<E::neq>(b.state, b.fs, idx, r);
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
}

// behind-side faces: lower faces for the forward sweep (geoLo), upper faces for the backward one
static __global__ void WaveGeoKernel(BlockDev b, PencilLattice L, int isViscous,
                                     double *__restrict__ geoLo, double *__restrict__ geoHi) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double lo[18], hi[18];
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
  lo[15] = hi[15] = lo[16] = hi[16] = lo[17] = hi[17] = 0.0;
  const long long t = PencilSlot(L, i, j, k);
  double2 *oLo = reinterpret_cast<double2 *>(geoLo + t * 18);
  double2 *oHi = reinterpret_cast<double2 *>(geoHi + t * 18);
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    oLo[q] = make_double2(lo[2 * q], lo[2 * q + 1]);
    oHi[q] = make_double2(hi[2 * q], hi[2 * q + 1]);
  }
}

template <int NS, int NT>
__global__ void __launch_bounds__(128)
    WaveDynKernel(BlockDev b, Params p, PencilLattice L, double *__restrict__ dyn) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT>;
  constexpr int neq = E::neq;
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double r[R::DN];
  MakeHead<NS, NT>(b, p, idx, r);
#pragma unroll
  for (int e = 0; e < neq; ++e) r[R::iB + e] = __ldg(b.rhs + e * b.fs + idx);
  r[R::iD] = __ldg(b.dinv + idx);
  r[R::iD + 1] = NT > 0 ? __ldg(b.dinv + b.fs + idx) : 0.0;
#pragma unroll
  for (int e = R::nUsed; e < R::DN; ++e) r[e] = 0.0;
  double2 *o = reinterpret_cast<double2 *>(dyn + PencilSlot(L, i, j, k) * R::DN);
#pragma unroll
  for (int q = 0; q < R::DN / 2; ++q) o[q] = make_double2(r[2 * q], r[2 * q + 1]);
}

// ---------------------------------------------------------------------------------------------
// ((0 + od_i) + od_j) + od_k over the three neighbours AHEAD of every cell, with the update as it
// is before the sweep (U of the forward sweep, L of the backward sweep; ref
// src/procBlock.cpp:1056-1170). Fully parallel, all reads from the block's fields (coalesced).
template <int NS, int NT, bool FORWARD>
__global__ void __launch_bounds__(128)
    LusgsAheadKernel(BlockDev b, Params p, PencilLattice L, double *__restrict__ ahead) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT>;
  constexpr int neq = E::neq, AN = R::AN;
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const int c[3] = {i, j, k}, nd[3] = {b.ni, b.nj, b.nk};
  const long long idx = CellIdx(b, i, j, k);
  double acc[AN];
#pragma unroll
  for (int e = 0; e < AN; ++e) acc[e] = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    const bool use = FORWARD ? (c[d] < nd[d] - 1 || ConnAcross(b, 2 * d + 2, c[d1], nd[d1], c[d2]))
                             : (c[d] > 0 || ConnAcross(b, 2 * d + 1, c[d1], nd[d1], c[d2]));
    if (!use) continue;
    const long long st = Stride(b, d);
    const long long idxn = FORWARD ? idx + st : idx - st;
    const long long fidx = FORWARD ? idx + st : idx;  // the face between the two cells
    double hd[R::NST], g[4], du[neq], sn[neq], Hn;
    MakeHead<NS, NT>(b, p, idxn, hd);
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = __ldg(b.fA[d] + q * b.fs + fidx);
    const double len = p.isViscous ? g[3] / __ldg(b.dist[d] + fidx) : 0.0;
#pragma unroll
    for (int e = 0; e < neq; ++e) du[e] = b.x[e * b.fs + idxn];
    MakeIngrDyn<NS, NT>(p.gas, hd, du, sn, &Hn);
    auto ld = [&](int cc) {
      return cc < neq + 2 ? hd[cc]
                          : (cc < 2 * neq + 2 ? du[cc - neq - 2]
                                              : (cc < 3 * neq + 2 ? sn[cc - 2 * neq - 2] : Hn));
    };
    // the ahead-neighbour is the geometrically upper one in a forward sweep
    OffDiagFromIngr<NS, NT>(ld, g, !FORWARD, acc, len * hd[R::iVt], len * hd[R::iVtT]);
  }
  double2 *o = reinterpret_cast<double2 *>(ahead + PencilSlot(L, i, j, k) * AN);
#pragma unroll
  for (int q = 0; q < AN / 2; ++q) o[q] = make_double2(acc[2 * q], acc[2 * q + 1]);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void BulkLoad(void *smemDst, const void *gsrc, unsigned bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(SmemAddr(smemDst)),
      "l"(gsrc), "r"(bytes), "r"(SmemAddr(bar))
      : "memory");
}

template <int NS, int NT>
struct PencilCfg {
  using R = PencilRec<NS, NT>;
  static constexpr int neq = NS + 4 + NT;
  static constexpr int NCOMP = kPCells * 4;   // four lanes per cell
  static constexpr int threads = NCOMP + 32;  // + the service warp (poller, publisher, loader)
  static constexpr int S = 6;                 // ring of plane stages: copies run S - 2 planes ahead
  static constexpr int dynB = kPCells * R::DN * 8, geoB = kPCells * R::GN * 8, ahB = kPCells * R::AN * 8;
  static constexpr int stageB = dynB + geoB + ahB;
  static_assert(dynB % 16 == 0 && geoB % 16 == 0 && ahB % 16 == 0, "bulk copies move 16-byte words");
  static constexpr int sxB = 2 * neq * kPCells * 8;
  static constexpr size_t smemBytes = static_cast<size_t>(S) * stageB + sxB + 8 * S + 16;
};

// what a lane fetches from global memory one plane ahead of its use: only lanes whose
// behind-neighbour is NOT in the pencil (cell of the pencil behind, ghost cell across a connection)
template <int NS, int NT>
struct PencilFetch {
  static constexpr int neq = NS + 4 + NT;
  double hd[PencilRec<NS, NT>::NST];  // record head of that neighbour
  double du[neq];                     // its update
};

// A cell (sweep coordinate I = q - jl - kl at plane q of its line) has FOUR lanes. Lane d = 0, 1, 2
// owns the behind-neighbour in direction i, j, k: it takes that neighbour's NEW update (shared
// memory, written one plane ago) and record head (the previous plane's stage), forms its new
// ingredients and the product with the cell's own face; the three products are summed in the
// reference's order with shuffles, the precomputed ahead-sum is added and lane l finishes
// equations l, l + 4, ... . The chain of a plane is
//     shared memory -> ingredients -> product -> shuffles -> solve -> shared memory.
template <int NS, int NT, bool FORWARD>
__global__ void __launch_bounds__(PencilCfg<NS, NT>::threads, kPencilCtasPerSm)
    LusgsPencilKernel(BlockDev b, Params p, PencilLattice L, int fullGS,
                      const double *__restrict__ dyn, const double *__restrict__ geo,
                      const double *__restrict__ ahead, const int2 *__restrict__ order,
                      int nPencils, WaveSync *sync, long long *dbg = nullptr) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT>;
  using C = PencilCfg<NS, NT>;
  using F = PencilFetch<NS, NT>;
  constexpr int neq = E::neq, nf = NS + 4, S = C::S;
  constexpr int TJ = kPTJ, TK = kPTK, NCOMP = C::NCOMP;
  constexpr int kPublish = 4;  // progress is announced every 4th plane (st.release ~1 000 cycles)
  constexpr int NOWN = (neq + 3) / 4;  // equations finished by one lane
  extern __shared__ __align__(128) unsigned char smemRaw[];
  auto stDyn = [&](int s) { return reinterpret_cast<const double *>(smemRaw + s * C::stageB); };
  auto stGeo = [&](int s) {
    return reinterpret_cast<const double *>(smemRaw + s * C::stageB + C::dynB);
  };
  auto stAh = [&](int s) {
    return reinterpret_cast<const double *>(smemRaw + s * C::stageB + C::dynB + C::geoB);
  };
  double *sxBase = reinterpret_cast<double *>(smemRaw + S * C::stageB);
  auto sx = [&](int par, int e, int cell) -> double & {
    return sxBase[(par * neq + e) * kPCells + cell];
  };
  uint64_t *full = reinterpret_cast<uint64_t *>(smemRaw + S * C::stageB + C::sxB);
  __shared__ int sTicket;

  const int tid = threadIdx.x;
  const bool isService = tid >= NCOMP;
  const int lane = tid & 3, cellS = isService ? 0 : tid >> 2;  // cell in sweep space
  const int jlS = cellS % TJ, klS = cellS / TJ;
  const int d = lane % 3;  // lane 3 has no neighbour (it mirrors lane 0's direction, unused)
  const int base = (tid & 31) & ~3;
  const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
  const int nd[3] = {b.ni, b.nj, b.nk};
  const long long strideD = Stride(b, d);
  int *done = sync->done;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) MbarInit(full + s, 1);
    MbarInitFence();
  }
  // planes handled by this thread block so far: plane q of the current pencil lives in stage
  // (fills + q) % S and completes phase ((fills + q) / S) & 1 of that stage's barrier
  int fills = 0;

  for (;;) {
    if (tid == 0) sTicket = static_cast<int>(atomicAdd(&sync->ticket, 1u));
    __syncthreads();
    const int ticket = sTicket;
    if (ticket >= nPencils) return;
    const int2 bc = order[ticket];  // sweep-space pencil
    const int bJ = FORWARD ? bc.x : L.nbJ - 1 - bc.x, bK = FORWARD ? bc.y : L.nbK - 1 - bc.y;
    const int j0 = bJ * TJ, k0 = bK * TK;
    const int tj = min(TJ, b.nj - j0), tk = min(TK, b.nk - k0);
    const int nSteps = b.ni + tj + tk - 2;
    const long long planeBase = static_cast<long long>(bJ + L.nbJ * bK) * L.planesPer;
    // sweep plane q -> plane of the workspace (geometric numbering)
    auto planeOf = [&](int q) { return FORWARD ? q : nSteps - 1 - q; };

    if (isService) {
      // ---- service warp. Lane 0 waits for the two pencils behind, lane 1 announces this pencil's
      // progress (two threads, because st.release is a gpu-scope fence: issued by the polling
      // thread it sat in series with the poll), lane 2 keeps the ring of plane stages filled.
      int *myFlag = done + bc.x + L.nbJ * bc.y;
      const int *flagJ = bc.x > 0 ? done + (bc.x - 1) + L.nbJ * bc.y : nullptr;
      const int *flagK = bc.y > 0 ? done + bc.x + L.nbJ * (bc.y - 1) : nullptr;
      // extents of the pencils behind (the clipped pencil is the first one of a backward sweep)
      const int tjB = FORWARD ? TJ : min(TJ, b.nj - (bJ + 1) * TJ);
      const int tkB = FORWARD ? TK : min(TK, b.nk - (bK + 1) * TK);
      int seenJ = 0, seenK = 0;
      // the foreign cell read at local plane q lies on plane q + tjB - 1 (k: q + tkB - 1) of the
      // pencil behind
      auto waitFor = [&](int q) {
        if (flagJ)
          while (seenJ < q + tjB) seenJ = LdAcquire(flagJ);
        if (flagK)
          while (seenK < q + tkB) seenK = LdAcquire(flagK);
      };
      auto load = [&](int q) {  // plane q of this pencil into its stage
        if (q >= nSteps) return;
        const int s = (fills + q) % S;
        const long long slot0 = (planeBase + planeOf(q)) * kPCells;
        uint64_t *bar = full + s;
        MbarExpectTx(bar, C::dynB + C::geoB + (fullGS ? C::ahB : 0));
        BulkLoad(smemRaw + s * C::stageB, dyn + slot0 * R::DN, C::dynB, bar);
        BulkLoad(smemRaw + s * C::stageB + C::dynB, geo + slot0 * R::GN, C::geoB, bar);
        if (fullGS)
          BulkLoad(smemRaw + s * C::stageB + C::dynB + C::geoB, ahead + slot0 * R::AN, C::ahB, bar);
      };
      if (tid == NCOMP + 2)
        for (int q = 0; q < S - 2; ++q) load(q);
      if (tid == NCOMP) waitFor(1);
      __syncthreads();
      for (int q = 0; q < nSteps; ++q) {
        // the stage of plane q + S - 2 held plane q - 2: its last readers finished with plane q - 1
        if (tid == NCOMP + 2) load(q + S - 2);
        // the lanes fetch the update of their NEXT plane's foreign neighbour at the top of a plane
        if (tid == NCOMP) waitFor(q + 2);
        // planes 0 .. q-1 are in global memory
        if (tid == NCOMP + 1 && q > 0 && (q % kPublish) == 0) StRelease(myFlag, q);
        __syncthreads();
      }
      if (tid == NCOMP + 1) StRelease(myFlag, kWaveDone);
      fills += nSteps;
      continue;
    }

    // ---- compute threads ---------------------------------------------------------------------
    // sweep-local line (jlS, klS) -> line of the block; the workspace numbers cells geometrically
    const bool lineValid = jlS < tj && klS < tk;
    const int jl = FORWARD ? jlS : tj - 1 - jlS, kl = FORWARD ? klS : tk - 1 - klS;
    const int j = j0 + jl, k = k0 + kl;
    const int cellG = lineValid ? jl + TJ * kl : 0;
    // the behind-neighbour in this lane's direction: cell of the workspace plane, inside the pencil?
    const int nbCellG = d == 0 ? cellG
                               : (d == 1 ? (FORWARD ? cellG - 1 : cellG + 1)
                                         : (FORWARD ? cellG - TJ : cellG + TJ));
    const bool nbInside = d == 0 ? true : (d == 1 ? jlS > 0 : klS > 0);
    const long long idxRow = lineValid ? CellIdx(b, 0, j, k) : 0;
    // sweep coordinate I -> cell index along the line
    auto iOf = [&](int I) { return FORWARD ? I : b.ni - 1 - I; };

    // Does the behind-neighbour of the cell solved at plane q contribute (physical cell, or across
    // a connection: ref src/procBlock.cpp:1064,1115; behind = lower side in a forward sweep), and
    // is it a cell of this pencil (else: record head formed on the fly, update read at L2)?
    // Along a line both answers are constants except at the line's first cell (direction i) and
    // on lines next to a block face with connection patches (the mask varies along i).
    const bool lineActive = lineValid && lane != 3;
    const bool onFace = d == 0 ? false : (FORWARD ? (d == 1 ? j == 0 : k == 0)
                                                  : (d == 1 ? j == b.nj - 1 : k == b.nk - 1));
    const bool faceHasConn = onFace && b.connFace[FORWARD ? 2 * d : 2 * d + 1] != nullptr;
    auto classify = [&](int q, bool *use, bool *inside) {
      const int I = q - jlS - klS;
      *use = false;
      *inside = false;
      if (!lineActive || I < 0 || I >= b.ni) return;
      if (d == 0) {
        *inside = I > 0;
        *use = I > 0 || ConnAcross(b, FORWARD ? 1 : 2, j, b.nj, k);
      } else {
        *inside = nbInside;
        if (!onFace) *use = true;
        else if (faceHasConn) {
          const int c[3] = {iOf(I), j, k};
          *use = ConnAcross(b, FORWARD ? 2 * d + 1 : 2 * d + 2, c[d1], nd[d1], c[d2]);
        }
      }
    };
    auto fetch = [&](int q, F &f) {
      bool use, inside;
      classify(q, &use, &inside);
      if (!use || inside) return;
      const long long idx = idxRow + iOf(q - jlS - klS);
      const long long nidx = FORWARD ? idx - strideD : idx + strideD;
      MakeHead<NS, NT>(b, p, nidx, f.hd);
      // updates of other pencils are rewritten during the sweep: L2, never L1
#pragma unroll
      for (int e = 0; e < neq; ++e) f.du[e] = __ldcg(b.x + e * b.fs + nidx);
    };

    auto step = [&](int q, const F &f, F &fNext) {
      fetch(q + 1, fNext);  // in flight during this plane
      const int I = q - jlS - klS;
      const bool active = lineValid && I >= 0 && I < b.ni;
      bool use, inside;
      classify(q, &use, &inside);
      const int g = fills + q;
      const int s = g % S, sPrev = (g + S - 1) % S;
      const bool rec_ = dbg != nullptr && blockIdx.x == 0 && tid == 0 && fills == 0 && q >= 64 && q < 96;
      if (rec_) dbg[(q - 64) * 8 + 0] = clock64();
      MbarWait(full + s, (g / S) & 1);
      if (rec_) dbg[(q - 64) * 8 + 1] = clock64();
      const double *myDyn = stDyn(s) + cellG * R::DN;
      // this lane's equations: right-hand side, D^-1, ahead-sum -- read before the product starts
      double ownB[NOWN], ownD[NOWN], ownA[NOWN];
#pragma unroll
      for (int hh = 0; hh < NOWN; ++hh) {
        const int e = lane + 4 * hh;
        if (e < neq) {
          ownB[hh] = myDyn[R::iB + e];
          ownD[hh] = myDyn[R::iD + (e < nf ? 0 : 1)];
          ownA[hh] = fullGS ? stAh(s)[cellG * R::AN + e] : 0.0;
        }
      }
      double od[neq];
#pragma unroll
      for (int e = 0; e < neq; ++e) od[e] = 0.0;
      if (use) {
        double hd[R::NST], du[neq], sn[neq], Hn;
        if (inside) {
          const double *nb = stDyn(sPrev) + nbCellG * R::DN;
#pragma unroll
          for (int e = 0; e < R::NST; ++e) hd[e] = nb[e];
#pragma unroll
          for (int e = 0; e < neq; ++e) du[e] = sx((q + 1) & 1, e, nbCellG);
        } else {
#pragma unroll
          for (int e = 0; e < R::NST; ++e) hd[e] = f.hd[e];
#pragma unroll
          for (int e = 0; e < neq; ++e) du[e] = f.du[e];
        }
        const double *gg = stGeo(s) + cellG * R::GN;
        double fa[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) fa[qq] = gg[4 * d + qq];
        const double len = gg[12 + d];
        MakeIngrDyn<NS, NT>(p.gas, hd, du, sn, &Hn);
        auto ld = [&](int cc) {
          return cc < neq + 2 ? hd[cc]
                              : (cc < 2 * neq + 2 ? du[cc - neq - 2]
                                                  : (cc < 3 * neq + 2 ? sn[cc - 2 * neq - 2] : Hn));
        };
        OffDiagFromIngr<NS, NT>(ld, fa, FORWARD, od, len * hd[R::iVt], len * hd[R::iVtT]);
      }
      if (rec_) dbg[(q - 64) * 8 + 2] = clock64() + (od[0] == 1.2345e300);
      // behind-sum (od_i + od_j) + od_k in every lane by a butterfly: lanes (0,1) and (2,3) swap,
      // then the pairs swap; lane 3 holds zero and addition commutes, so each lane forms exactly
      // the reference's ((0 + od_i) + od_j) + od_k. Then this lane's equations.
      // forward: x = D^-1 (b + (L - U)); backward: D^-1 ((b + L) - U), or on the first sweep
      // without initialisation x - D^-1 U (ref src/linearSolver.cpp:341-428)
#pragma unroll
      for (int e = 0; e < neq; ++e) {
        const double pr = od[e] + __shfl_xor_sync(0xffffffffu, od[e], 1);
        const double other = __shfl_xor_sync(0xffffffffu, pr, 2);
        const double bs = (lane & 2) ? other + pr : pr + other;
        if ((e & 3) == lane && active) {
          const double as = fullGS ? ownA[e >> 2] : 0.0;
          const double rb = ownB[e >> 2];
          double r;
          if (FORWARD) r = rb + (bs - as);
          else if (fullGS) r = (rb + as) - bs;
          else r = bs;
          r *= ownD[e >> 2];
          const long long gi = e * b.fs + idxRow + iOf(I);
          if (!FORWARD && !fullGS) r = b.x[gi] - r;
          sx(q & 1, e, cellG) = r;
          __stcg(b.x + gi, r);
        }
      }
      if (rec_) dbg[(q - 64) * 8 + 3] = clock64();
      __syncthreads();
      if (rec_) dbg[(q - 64) * 8 + 4] = clock64();
    };

    __syncthreads();  // the poller has seen what planes 0 and 1 read
    F fa_, fb_;
    fetch(0, fa_);
    for (int q = 0; q < nSteps; q += 2) {
      step(q, fa_, fb_);
      if (q + 1 < nSteps) step(q + 1, fb_, fa_);
    }
    fills += nSteps;
  }
}

}  // namespace aither
