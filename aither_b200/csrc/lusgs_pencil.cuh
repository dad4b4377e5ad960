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
//      the first half sweep of an iteration, and every half sweep leaves the next one's sums behind
//      (its behind-sums with the new update ARE the next sweep's ahead-sums with the old one);
//   3. the wavefront proper (LusgsPencilKernel): the block is cut into pencils of 12 x 8 grid
//      lines; a thread block walks its pencil in local planes q = i + jl + kl, one barrier per
//      plane, one thread per cell, the new ingredients handed on through shared memory; pencils
//      are ordered by atomic tickets along anti-diagonals, so a predecessor is always running or
//      done, and take the boundary lines of the two pencils behind them from a mailbox of tagged
//      16-byte entries (no fences, no progress counters);
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

// pencil cross-section (tunable at build time for A/B runs, scripts/build_pencil_variants.sh):
// 12 x 8 = 96 grid lines, one thread each, + four service warps = 224 threads, one thread block
// per SM. Measured with the one-thread-per-cell kernel (profiles/r02ad_pencil_cross_section.json;
// LU-SGS x4 at 192^3 / SST + LU-SGS at 128^3, ms per iteration): 16x8 10.53 / 9.80, 12x8 10.27 /
// 9.71, 16x6 10.22, 8x12 10.18, 8x8 10.84 / 9.21.
#ifndef AITHER_PENCIL_TJ
#define AITHER_PENCIL_TJ 12
#endif
#ifndef AITHER_PENCIL_TK
#define AITHER_PENCIL_TK 8
#endif
#ifndef AITHER_PENCIL_CTAS
#define AITHER_PENCIL_CTAS 1
#endif
constexpr int kPTJ = AITHER_PENCIL_TJ, kPTK = AITHER_PENCIL_TK, kPCells = kPTJ * kPTK;
constexpr int kPencilCtasPerSm = AITHER_PENCIL_CTAS;

// VISC = false (Euler runs): the records carry no viscous slots -- 14 + 14 doubles per cell and
// plane instead of 18 + 18, a quarter of the workspace's bytes less and a deeper stage ring.
template <int NS, int NT, bool VISC = true>
struct PencilRec {
  static constexpr int neq = NS + 4 + NT;
  // per-iteration record: s[neq] | H a (vt vtT) | b[neq] | dinv dinvT, padded to a stride that is a
  // multiple of 16 bytes and NOT of 128 (records of neighbouring cells in different banks)
  static constexpr int NST = neq + (VISC ? 4 : 2);  // record head: s | H a (| vt vtT)
  static constexpr int nUsed = NST + neq + 2;
  static constexpr int nEven = (nUsed + 1) & ~1;
  static constexpr int DN = nEven % 16 == 0 ? nEven + 2 : nEven;
  static constexpr int iH = neq, iA = neq + 1, iVt = neq + 2, iVtT = neq + 3, iB = NST,
                       iD = NST + neq;
  // per-block record: behind-side faces i, j, k {nx, ny, nz, |A|} (| |A| / dist for i, j, k) | pad
  static constexpr int GN = VISC ? 18 : 14;
  static constexpr int AN = neq;  // ahead-sum: component-major within a plane, [plane][e][cell]
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

// record head of a cell: r[0, neq) holds its state on entry; what WaveDynKernel packs for the
// cells of the block and the halo warp forms on the fly (same expressions) for ghost cells and
// cells of other pencils
template <int NS, int NT, bool VISC>
__device__ __forceinline__ void HeadFromState(const Params &p, double *r, double mu, double mut,
                                              double f1) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT, VISC>;
  const MixK<NS> m = MixOf<NS>(p.gas, r);
  const double t0 = r[E::ie] * m.tFac;
  r[R::iH] = m.hf + m.cp * t0 + 0.5 * VelMagSq<NS>(r);  // as MakeIngr
  r[R::iA] = sqrt(m.gamma * r[E::ie] * m.rhoInv);
  if constexpr (VISC) {
    r[R::iVt] = 0.0;
    r[R::iVtT] = 0.0;
  }
  if (VISC && p.isViscous) {
    // state-dependent factors of the viscous face spectral radii (NeighbourViscTerms)
    const double rho = SpeciesSum<NS>(r);
    r[R::iVt] = ViscSpecFactor(p.tr, rho, Gamma<NS>(p.gas, r), mu, mut);
    if (NT > 0)
      r[R::iVtT] = TurbViscSpecFactor(p.tr.turbModel, p.tr.scaling, rho, r[NS + 4],
                                      r[NS + 4 + (NT > 1 ? 1 : 0)], mu, mut, f1);
  }
}
template <int NS, int NT, bool VISC>
__device__ __forceinline__ void MakeHead(const BlockDev &b, const Params &p, long long idx,
                                         double *r) {
  using E = Eq<NS, NT>;
  LoadCell<E::neq>(b.state, b.fs, idx, r);
  double mu = 0.0, mut = 0.0, f1 = 0.0;
  if (VISC && p.isViscous) {
    mu = __ldg(b.viscosity + idx);
    if (NT > 0) {
      mut = __ldg(b.eddyVisc + idx);
      f1 = __ldg(b.f1 + idx);
    }
  }
  HeadFromState<NS, NT, VISC>(p, r, mu, mut, f1);
}

// behind-side faces: lower faces for the forward sweep (geoLo), upper faces for the backward one
template <int GN>
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
    lo[12 + d] = (GN > 14 && isViscous) ? lo[4 * d + 3] / b.dist[d][idx] : 0.0;
    hi[12 + d] = (GN > 14 && isViscous) ? hi[4 * d + 3] / b.dist[d][idx + st] : 0.0;
  }
  lo[15] = hi[15] = lo[16] = hi[16] = lo[17] = hi[17] = 0.0;
  const long long t = PencilSlot(L, i, j, k);
  double2 *oLo = reinterpret_cast<double2 *>(geoLo + t * GN);
  double2 *oHi = reinterpret_cast<double2 *>(geoHi + t * GN);
#pragma unroll
  for (int q = 0; q < GN / 2; ++q) {
    oLo[q] = make_double2(lo[2 * q], lo[2 * q + 1]);
    oHi[q] = make_double2(hi[2 * q], hi[2 * q + 1]);
  }
}

template <int NS, int NT, bool VISC>
__global__ void __launch_bounds__(128)
    WaveDynKernel(BlockDev b, Params p, PencilLattice L, double *__restrict__ dyn) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT, VISC>;
  constexpr int neq = E::neq;
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double r[R::DN];
  MakeHead<NS, NT, VISC>(b, p, idx, r);
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
template <int NS, int NT, bool FORWARD, bool VISC>
__global__ void __launch_bounds__(128)
    LusgsAheadKernel(BlockDev b, Params p, PencilLattice L, double *__restrict__ ahead, int i0,
                     int j0, int k0, int i1, int j1) {
  // cells [i0, i1) x [j0, j1) x [k0, k0 + gridDim.z): the whole block, or the layer of cells next to
  // a connected face (whose ghost cells changed since the previous half sweep left its sums)
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT, VISC>;
  constexpr int neq = E::neq, AN = R::AN;
  const int i = i0 + blockIdx.x * 32 + threadIdx.x, j = j0 + blockIdx.y * 4 + threadIdx.y,
            k = k0 + blockIdx.z;
  if (i >= i1 || j >= j1) return;
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
    MakeHead<NS, NT, VISC>(b, p, idxn, hd);
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = __ldg(b.fA[d] + q * b.fs + fidx);
    const double len = (VISC && p.isViscous) ? g[3] / __ldg(b.dist[d] + fidx) : 0.0;
#pragma unroll
    for (int e = 0; e < neq; ++e) du[e] = b.x[e * b.fs + idxn];
    MakeIngrDyn<NS, NT>(p.gas, hd, du, sn, &Hn);
    auto ld = [&](int cc) {
      return cc < neq + 2 ? hd[cc]
                          : (cc < 2 * neq + 2 ? du[cc - neq - 2]
                                              : (cc < 3 * neq + 2 ? sn[cc - 2 * neq - 2] : Hn));
    };
    // the ahead-neighbour is the geometrically upper one in a forward sweep
    OffDiagFromIngr<NS, NT>(ld, g, !FORWARD, acc, VISC ? len * hd[VISC ? R::iVt : 0] : 0.0,
                            VISC ? len * hd[VISC ? R::iVtT : 0] : 0.0);
  }
  const long long t = PencilSlot(L, i, j, k);
  double *o = ahead + (t / kPCells) * (AN * kPCells) + t % kPCells;
#pragma unroll
  for (int e = 0; e < AN; ++e) o[e * kPCells] = acc[e];
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

template <int NS, int NT, bool VISC = true>
struct PencilCfg {
  using R = PencilRec<NS, NT, VISC>;
  static constexpr int neq = NS + 4 + NT;
  static constexpr int NCOMP = kPCells;  // one thread per grid line of the pencil
  static_assert(NCOMP % 32 == 0, "the walkers fill whole warps");
  // + four service warps taking turns as loader (ring of plane stages), mailbox writer (boundary
  // lines for the pencils ahead) and halo (ingredients of the neighbours outside the pencil)
  static constexpr int threads = NCOMP + 128;
  static constexpr int NH = kPTJ + 2 * kPTK;  // halo lanes: column behind in j, row behind in k, line starts
  static_assert(NH <= 32, "one warp prepares the halo of a plane");
  static constexpr int dynB = kPCells * R::DN * 8, geoB = kPCells * R::GN * 8, ahB = kPCells * R::AN * 8;
  static constexpr int stageB = dynB + geoB + ahB;
  static_assert(dynB % 16 == 0 && geoB % 16 == 0 && ahB % 16 == 0, "bulk copies move 16-byte words");
  // update-dependent ingredients of the cells solved one plane ago: du | sn | Hn, component-major
  // over the cross-section padded by one line behind in j and in k (filled by the halo warp)
  static constexpr int PJ = kPTJ + 1, PCELLS = PJ * (kPTK + 1);
  static constexpr int NI = 2 * neq + 1;
  static constexpr int ingB = 2 * NI * PCELLS * 8;
  static constexpr int haloB = 2 * NH * R::NST * 8;  // record heads of the halo cells
  static constexpr int fixedB = ingB + haloB + 8 * 8 + 16;
  static constexpr int kMaxSmem = 227 * 1024;
  static constexpr int sFit = (kMaxSmem - fixedB) / stageB;
  static constexpr int S = sFit > 6 ? 6 : sFit;  // ring of plane stages: copies run S - 2 planes ahead
  static_assert(S >= 3, "the cross-section does not fit shared memory");
  static constexpr size_t smemBytes = static_cast<size_t>(S) * stageB + fixedB;
};

__device__ __forceinline__ void NamedBarrier(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Hand-over between pencils without fences: a MAILBOX entry is 16 bytes, two 8-byte words
// {low half of the double, tag} {high half, tag}. Each 8-byte word is written atomically, so a
// reader that finds the expected tag in both words holds the value -- no release / acquire pair,
// no progress counter. (A gpu-scope release costs ~1 000 cycles on B200 and, announced every
// plane by every pencil, slowed the whole sweep: profiles/r02y, 2.0 -> 1.5 ms per half sweep just
// by announcing every 4th plane.) The tag is the number of the half sweep (never 0; the mailbox
// is zeroed when it is allocated), so what an earlier half sweep left behind never matches.
__device__ __forceinline__ void MailStore(uint4 *p, double v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "r"(static_cast<unsigned>(__double2loint(v))), "r"(tag),
               "r"(static_cast<unsigned>(__double2hiint(v))), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint4 MailLoad(const uint4 *p) {
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

// what a halo lane loads one plane ahead of its use. Neighbour in another pencil: record head from
// that pencil's workspace record, ingredients from the mailbox. Ghost cell: state (in hd), update
// (in the first neq entries) and viscosities from the block's fields.
template <int NS, int NT, bool VISC = true>
struct HaloRaw {
  static constexpr int neq = NS + 4 + NT;
  double hd[PencilRec<NS, NT, VISC>::NST];
  uint4 ent[2 * neq + 1];
  double mu, mut, f1;
  const uint4 *mail;  // where the entries come from (re-read until their tags match)
  int kind;           // 0 nothing, 1 mailbox, 2 ghost cell
};

// ONE thread per cell. A cell (sweep coordinate I = q - jl - kl at plane q of its line) takes the
// ingredients of its three behind-neighbours -- the one in i from its own registers (the same
// thread solved it one plane ago), the ones in j and k from shared memory, where their threads
// left them --, forms the three products side by side, sums them in the reference's order
// ((0 + od_i) + od_j) + od_k, adds the precomputed ahead-sum, solves its equations, and leaves its
// own new ingredients (MakeIngrDyn, once per cell instead of once per neighbour) for the next
// plane. The chain of a plane is
//     shared memory -> three products -> solve -> own ingredients -> shared memory -> barrier.
// Neighbours that are not cells of the pencil are served by the HALO warp, one lane per line on
// the pencil's two behind-sides: cells of the pencil behind arrive through the mailbox (written
// by that pencil's boundary lines together with their shared-memory copy), ghost cells across a
// connection and in front of a line are formed from the block's fields. The halo warp puts them
// where the walkers look (cross-section padded by one line), so the walkers run one uniform path.
// Pencils are handed out by an atomic ticket in anti-diagonal order: a pencil's predecessors are
// always running or done, whatever the number of resident thread blocks.
//
// History (192^3, x4 sweeps, wavefront ms per iteration; profiles/r02*): four lanes per cell,
// each forming its neighbour's ingredients, 8 x 7 / 8 x 8 cells: 18.4 / 17.4 -- fp64 ISSUE bound
// (a warp instruction costs the same with 8 cells in it as with 32); one thread per cell with
// progress counters (poller + publisher warp): 15.6, the pencils waiting on each other's
// st.release.
template <int NS, int NT, bool FORWARD, bool VISC>
__global__ void __launch_bounds__(PencilCfg<NS, NT, VISC>::threads, 1)
    LusgsPencilKernel(BlockDev b, Params p, PencilLattice L, int fullGS,
                      const double *__restrict__ dyn, const double *__restrict__ geo,
                      const double *__restrict__ ahead, const int2 *__restrict__ order,
                      int nPencils, WaveSync *sync, uint4 *mailJ, uint4 *mailK, unsigned tag,
                      double *__restrict__ carry, long long *dbg = nullptr, int dbgFlags = 0) {
  using E = Eq<NS, NT>;
  using R = PencilRec<NS, NT, VISC>;
  using C = PencilCfg<NS, NT, VISC>;
  using HR = HaloRaw<NS, NT, VISC>;
  constexpr int neq = E::neq, nf = NS + 4, S = C::S;
  constexpr int TJ = kPTJ, TK = kPTK, NCOMP = C::NCOMP, PJ = C::PJ, PCELLS = C::PCELLS, NI = C::NI;
  constexpr int NH = C::NH, NST = R::NST;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  auto stDyn = [&](int s) { return reinterpret_cast<const double *>(smemRaw + s * C::stageB); };
  auto stGeo = [&](int s) {
    return reinterpret_cast<const double *>(smemRaw + s * C::stageB + C::dynB);
  };
  auto stAh = [&](int s) {
    return reinterpret_cast<const double *>(smemRaw + s * C::stageB + C::dynB + C::geoB);
  };
  double *ingBase = reinterpret_cast<double *>(smemRaw + S * C::stageB);
  auto ing = [&](int par, int e, int P) -> double & { return ingBase[(par * NI + e) * PCELLS + P]; };
  double *haloBase = reinterpret_cast<double *>(smemRaw + S * C::stageB + C::ingB);
  auto haloHd = [&](int par, int h) { return haloBase + (par * NH + h) * NST; };
  uint64_t *full = reinterpret_cast<uint64_t *>(smemRaw + S * C::stageB + C::ingB + C::haloB);
  __shared__ int sTicket;

  const int tid = threadIdx.x;
  const int role = tid < NCOMP ? 0 : (tid - NCOMP) / 32 + 1;  // 0 walker, 1 .. 4 service
  const int nd[3] = {b.ni, b.nj, b.nk};

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
    // sweep coordinate I -> cell index along the line
    auto iOf = [&](int I) { return FORWARD ? I : b.ni - 1 - I; };
    // mailboxes are numbered in sweep space: [pencil][plane of the READER][line][entry]
    const long long myMail = static_cast<long long>(bc.x + L.nbJ * bc.y) * L.planesPer;
    if (role != 0) {
      // ---- four service warps, taking turns: warp w serves the planes p = w (mod 4). During
      // plane p - 2 it loads what the halo of plane p needs, during plane p - 1 it puts that halo in
      // place; during plane p - 3 it is the loader (ring of plane stages) and hands the boundary
      // lines of the plane just finished to the pencils ahead. Each of these is a few hundred
      // dependent instructions of ONE warp -- about as long as the walkers' plane; taking turns
      // keeps them off the plane's critical path (one halo warp: 0.94 -> 1.4 us per plane for a
      // pencil with both neighbours behind, profiles/r02y).
      const int w = role - 1;
      const int h = tid & 31;
      // -- loader
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
      // -- mailbox out: lane m < TK posts line (tj - 1, m) for the pencil ahead in j, lane
      // TK <= m < TK + TJ line (m - TK, tk - 1) for the one ahead in k. Plane q of this pencil is
      // plane q - (tj - 1) of the reader in j (its line 0 lies tj - 1 planes behind this pencil's
      // last line), q - (tk - 1) in k.
      const bool toJ = h < TK;
      const int pjS = toJ ? tj - 1 : h - TK, pkS = toJ ? h : tk - 1;
      const bool posts = h < TK + TJ && pjS < tj && pkS < tk &&
                         (toJ ? bc.x + 1 < L.nbJ : bc.y + 1 < L.nbK);
      const int postP = (pjS + 1) + PJ * (pkS + 1);
      // a mailbox plane is entry-major, [entry][line]: the lanes of one load / store instruction
      // touch one or two 128-byte lines (line-major, every lane its own line: the load / store
      // unit took them one line per cycle, in the way of the walkers' shared-memory traffic)
      const int postLines = toJ ? TK : TJ;
      uint4 *out = toJ ? mailJ + (myMail + L.planesPer - (tj - 1)) * (TK * NI) + pkS
                       : mailK + (myMail + static_cast<long long>(L.nbJ) * L.planesPer - (tk - 1)) * (TJ * NI) + pjS;
      auto post = [&](int q) {
        const int I = q - pjS - pkS;
        if (!posts || I < 0 || I >= b.ni || (dbgFlags & 2)) return;
        uint4 *o = out + static_cast<long long>(q) * (postLines * NI);
#pragma unroll
        for (int e = 0; e < NI; ++e) MailStore(o + e * postLines, ing(q & 1, e, postP), tag);
      };
      // -- halo in. Lane h < TK: the cell behind (in j) line (0, h); TK <= h < TK + TJ: the cell
      // behind (in k) line (h - TK, 0); then TK lanes for the ghost cell in front of the line that
      // starts at this plane (jl = q - kl, kl = h - TK - TJ)
      const int hd = h < TK ? 1 : (h < TK + TJ ? 2 : 0);
      const long long strideH = Stride(b, hd);
      const int hd1 = (hd + 1) % 3, hd2 = (hd + 2) % 3;
      // sweep-space line of the cell this lane serves at plane q
      auto lineOf = [&](int q, int *jS, int *kS) {
        if (hd == 1) {
          *jS = 0;
          *kS = h;
        } else if (hd == 2) {
          *jS = h - TK;
          *kS = 0;
        } else {
          *kS = h - TK - TJ;
          *jS = q - *kS;
        }
        return h < NH && q < nSteps && *jS >= 0 && *jS < tj && *kS < tk;
      };
      // neighbour in the pencil behind (not at the block's face): mailbox of this pencil
      const bool fromPencil = hd == 1 ? bc.x > 0 : (hd == 2 ? bc.y > 0 : false);
      const int mailLines = hd == 1 ? TK : TJ;
      auto hfetch = [&](int q, HR &f) {
        f.kind = 0;
        int jS, kS;
        if (!lineOf(q, &jS, &kS)) return;
        const int I = q - jS - kS;
        if (I < 0 || I >= b.ni) return;
        const int j = j0 + (FORWARD ? jS : tj - 1 - jS), k = k0 + (FORWARD ? kS : tk - 1 - kS);
        const int c[3] = {iOf(I), j, k};
        if (fromPencil && (dbgFlags & 1)) return;
        if (fromPencil) {
          int cn[3] = {c[0], c[1], c[2]};
          cn[hd] += FORWARD ? -1 : 1;
          const double2 *rec =
              reinterpret_cast<const double2 *>(dyn + PencilSlot(L, cn[0], cn[1], cn[2]) * R::DN);
#pragma unroll
          for (int e = 0; e < NST / 2; ++e) {
            const double2 v = __ldg(rec + e);
            f.hd[2 * e] = v.x;
            f.hd[2 * e + 1] = v.y;
          }
          if (NST & 1) f.hd[NST - 1] = __ldg(reinterpret_cast<const double *>(rec) + NST - 1);
          f.mail = hd == 1 ? mailJ + (myMail + q) * (TK * NI) + kS : mailK + (myMail + q) * (TJ * NI) + jS;
#pragma unroll
          for (int e = 0; e < NI; ++e) f.ent[e] = MailLoad(f.mail + e * mailLines);
          f.kind = 1;
          return;
        }
        // ghost cell: does it contribute (across a connection; ref src/procBlock.cpp:1064,1115;
        // behind = lower side in a forward sweep)?
        if (!ConnAcross(b, FORWARD ? 2 * hd + 1 : 2 * hd + 2, c[hd1], nd[hd1], c[hd2])) return;
        const long long idx = CellIdx(b, c[0], j, k);
        const long long nidx = FORWARD ? idx - strideH : idx + strideH;
        LoadCell<neq>(b.state, b.fs, nidx, f.hd);
#pragma unroll
        for (int e = 0; e < neq; ++e) {
          const double v = __ldcg(b.x + e * b.fs + nidx);
          f.ent[e] = make_uint4(__double2loint(v), tag, __double2hiint(v), tag);
        }
        f.mu = f.mut = f.f1 = 0.0;
        if (VISC && p.isViscous) {
          f.mu = __ldg(b.viscosity + nidx);
          if (NT > 0) {
            f.mut = __ldg(b.eddyVisc + nidx);
            f.f1 = __ldg(b.f1 + nidx);
          }
        }
        f.kind = 2;
      };
      auto hstore = [&](int q, HR &f) {
        if (f.kind == 0) return;
        int jS, kS;
        lineOf(q, &jS, &kS);
        double v[NI];
        if (f.kind == 1) {
          // The pencil behind may not be there yet. Pencils run at the same pace, so a pencil that
          // has caught up waits here every plane: poll ONE entry, with a pause, and re-read the
          // rest only when it has arrived.
          for (;;) {
            unsigned bad = 0;
#pragma unroll
            for (int e = 0; e < NI; ++e) bad |= (f.ent[e].y ^ tag) | (f.ent[e].w ^ tag);
            if (bad == 0 || (dbgFlags & 4)) break;
            for (;;) {
              const uint4 t = MailLoad(f.mail + (NI - 1) * mailLines);
              if (t.y == tag && t.w == tag) break;
              __nanosleep(200);
            }
#pragma unroll
            for (int e = 0; e < NI; ++e) f.ent[e] = MailLoad(f.mail + e * mailLines);
          }
#pragma unroll
          for (int e = 0; e < NI; ++e) v[e] = __hiloint2double(f.ent[e].z, f.ent[e].x);
        } else {
#pragma unroll
          for (int e = 0; e < neq; ++e) v[e] = __hiloint2double(f.ent[e].z, f.ent[e].x);
          HeadFromState<NS, NT, VISC>(p, f.hd, f.mu, f.mut, f.f1);
          MakeIngrDyn<NS, NT>(p.gas, f.hd, v, v + neq, v + 2 * neq);
        }
        double *o = haloHd(q & 1, h);
#pragma unroll
        for (int e = 0; e < NST; ++e) o[e] = f.hd[e];
        // where the walker of plane q looks for this neighbour: one line behind in j / k, or (line
        // start) at its own place
        const int P = (hd == 1 ? 0 : jS + 1) + PJ * (hd == 2 ? 0 : kS + 1);
        const int par = (q + 1) & 1;
#pragma unroll
        for (int e = 0; e < NI; ++e) ing(par, e, P) = v[e];
      };
      HR hr;
      hr.kind = 0;
      if (w == 0) {
        hfetch(0, hr);
        hstore(0, hr);
      } else if (w == 1) {
        hfetch(1, hr);
      } else if (w == 3 && h == 0) {
        for (int q = 0; q < S - 2; ++q) load(q);
      }
      NamedBarrier(1, C::threads);
      for (int q = 0; q < nSteps; ++q) {
        // 0: idle, 1: load + post, 2: fetch plane q + 2, 3: halo of plane q + 1 in place. (Fetching
        // three planes ahead kept every pencil one plane further behind the pencils it follows.)
        const int turn = (q - w) & 3;
        if (turn == 2) {
          hfetch(q + 2, hr);
        } else if (turn == 1) {
          // the stage of plane q + S - 2 held plane q - 2: its last readers finished with plane q - 1
          if (h == 0) load(q + S - 2);
          if (q > 0) post(q - 1);
        } else if (turn == 3) {
          hstore(q + 1, hr);
        }
        NamedBarrier(1, C::threads);
      }
      if (w == 0) post(nSteps - 1);
      fills += nSteps;
      continue;
    }

    // ---- walkers -------------------------------------------------------------------------------
    const int jlS = tid % TJ, klS = tid / TJ;  // line in sweep space
    const int P = (jlS + 1) + PJ * (klS + 1);
    // sweep-local line (jlS, klS) -> line of the block; the workspace numbers cells geometrically
    const bool lineValid = jlS < tj && klS < tk;
    const int jl = FORWARD ? jlS : tj - 1 - jlS, kl = FORWARD ? klS : tk - 1 - klS;
    const int j = j0 + jl, k = k0 + kl;
    const int cellG = lineValid ? jl + TJ * kl : 0;
    const int nbG1 = FORWARD ? cellG - 1 : cellG + 1, nbG2 = FORWARD ? cellG - TJ : cellG + TJ;
    const long long idxRow = lineValid ? CellIdx(b, 0, j, k) : 0;
    // Does a behind-neighbour contribute (physical cell, or across a connection)? Along a line the
    // answer is constant except at the line's first cell (direction i) and on lines next to a
    // block face with connection patches (the mask varies along i).
    const bool use0Start = lineValid && ConnAcross(b, FORWARD ? 1 : 2, j, b.nj, k);
    const bool onFace1 = FORWARD ? j == 0 : j == b.nj - 1, onFace2 = FORWARD ? k == 0 : k == b.nk - 1;
    const bool conn1 = onFace1 && b.connFace[FORWARD ? 2 : 3] != nullptr;
    const bool conn2 = onFace2 && b.connFace[FORWARD ? 4 : 5] != nullptr;
    // ingredients of this line's previous cell (the behind-neighbour in i), carried in registers
    double pHd[NST], pDu[neq], pSn[neq], pHn = 0.0;
#pragma unroll
    for (int e = 0; e < NST; ++e) pHd[e] = 0.0;
#pragma unroll
    for (int e = 0; e < neq; ++e) pDu[e] = pSn[e] = 0.0;

    auto step = [&](int q) {
      const int I = q - jlS - klS;
      const bool active = lineValid && I >= 0 && I < b.ni;
      const int g = fills + q;
      const int s = g % S, sPrev = (g + S - 1) % S;
      MbarWait(full + s, (g / S) & 1);
      if (active) {
        const int ic = iOf(I);
        const double *myDyn = stDyn(s) + cellG * R::DN;
        const double *gg = stGeo(s) + cellG * R::GN;
        const int parR = (q + 1) & 1;
        if (I == 0) {  // line start: the ghost cell in front of the line, from the halo warp
          const double *hp = haloHd(q & 1, TK + TJ + klS);
#pragma unroll
          for (int e = 0; e < NST; ++e) pHd[e] = hp[e];
#pragma unroll
          for (int e = 0; e < neq; ++e) {
            pDu[e] = ing(parR, e, P);
            pSn[e] = ing(parR, neq + e, P);
          }
          pHn = ing(parR, 2 * neq, P);
        }
        // The three products are formed unconditionally and side by side (independent chains); a
        // neighbour that does not contribute (block face without connection) is dropped by a
        // select, so whatever its slot holds never reaches the sum.
        const bool use0 = I > 0 || use0Start;
        const bool use1 = !onFace1 || (conn1 && ConnAcross(b, FORWARD ? 3 : 4, k, b.nk, ic));
        const bool use2 = !onFace2 || (conn2 && ConnAcross(b, FORWARD ? 5 : 6, ic, b.ni, j));
        auto product = [&](int d, const double *hd, const double *du, const double *sn,
                           double Hn, double *od) {
          double fa[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) fa[qq] = gg[4 * d + qq];
          const double len = VISC ? gg[VISC ? 12 + d : 0] : 0.0;
          auto ld = [&](int cc) {
            return cc < neq + 2 ? hd[cc]
                                : (cc < 2 * neq + 2 ? du[cc - neq - 2]
                                                    : (cc < 3 * neq + 2 ? sn[cc - 2 * neq - 2] : Hn));
          };
#pragma unroll
          for (int e = 0; e < neq; ++e) od[e] = 0.0;
          OffDiagFromIngr<NS, NT>(ld, fa, FORWARD, od, VISC ? len * hd[VISC ? R::iVt : 0] : 0.0,
                                  VISC ? len * hd[VISC ? R::iVtT : 0] : 0.0);
        };
        auto fromSmem = [&](int d, const double *hdp, int Pn, double *od) {
          double hd[NST], du[neq], sn[neq];
#pragma unroll
          for (int e = 0; e < NST; ++e) hd[e] = hdp[e];
#pragma unroll
          for (int e = 0; e < neq; ++e) {
            du[e] = ing(parR, e, Pn);
            sn[e] = ing(parR, neq + e, Pn);
          }
          product(d, hd, du, sn, ing(parR, 2 * neq, Pn), od);
        };
        double od0[neq], od1[neq], od2[neq];
        product(0, pHd, pDu, pSn, pHn, od0);
        fromSmem(1, jlS > 0 ? stDyn(sPrev) + nbG1 * R::DN : haloHd(q & 1, klS), P - 1, od1);
        fromSmem(2, klS > 0 ? stDyn(sPrev) + nbG2 * R::DN : haloHd(q & 1, TK + jlS), P - PJ, od2);
        // forward: x = D^-1 (b + (L - U)); backward: D^-1 ((b + L) - U), or on the first sweep
        // without initialisation x - D^-1 U (ref src/linearSolver.cpp:341-428)
#pragma unroll
        for (int e = 0; e < NST; ++e) pHd[e] = myDyn[e];
        const double *ah = stAh(s) + cellG;
        double bsv[neq], xn[neq];
#pragma unroll
        for (int e = 0; e < neq; ++e) {
          double bs = 0.0;
          bs = use0 ? bs + od0[e] : bs;
          bs = use1 ? bs + od1[e] : bs;
          bs = use2 ? bs + od2[e] : bs;
          bsv[e] = bs;
          const double as = fullGS ? ah[e * kPCells] : 0.0;
          const double rb = myDyn[R::iB + e];
          double r;
          if (FORWARD) r = rb + (bs - as);
          else if (fullGS) r = (rb + as) - bs;
          else r = bs;
          r *= myDyn[R::iD + (e < nf ? 0 : 1)];
          if (!FORWARD && !fullGS) r = b.x[e * b.fs + idxRow + ic] - r;
          xn[e] = r;
        }
        // the chain first: own new ingredients, handed on through shared memory
        MakeIngrDyn<NS, NT>(p.gas, pHd, xn, pSn, &pHn);
        const int parW = q & 1;
#pragma unroll
        for (int e = 0; e < neq; ++e) {
          ing(parW, e, P) = xn[e];
          ing(parW, neq + e, P) = pSn[e];
        }
        ing(parW, 2 * neq, P) = pHn;
        // ... then what nobody waits for. The update goes to the block's field two cells of the
        // line at a time (16-byte stores, half the store instructions and sectors): pDu still
        // holds the update of the line's previous cell (ic - 1 forward, ic + 1 backward).
        const bool pairHi = (ic & 1) != 0;  // the pair is (ic - 1, ic) or (ic, ic + 1)
        const long long gi0 = idxRow + ic;
        if (FORWARD ? pairHi : (!pairHi && I > 0)) {
#pragma unroll
          for (int e = 0; e < neq; ++e) {
            if (FORWARD) __stcg(reinterpret_cast<double2 *>(b.x + e * b.fs + gi0 - 1), make_double2(pDu[e], xn[e]));
            else __stcg(reinterpret_cast<double2 *>(b.x + e * b.fs + gi0), make_double2(xn[e], pDu[e]));
          }
        } else if (FORWARD ? ic == b.ni - 1 : !pairHi) {
#pragma unroll
          for (int e = 0; e < neq; ++e) __stcg(b.x + e * b.fs + gi0, xn[e]);
        }
#pragma unroll
        for (int e = 0; e < neq; ++e) pDu[e] = xn[e];
        // The sum over this sweep's behind-neighbours with their NEW update is the next half
        // sweep's sum over its ahead-neighbours with their OLD update (same neighbours, same
        // update, same faces): it goes where this plane's ahead-sum came from, and the parallel
        // ahead-sum pass (LusgsAheadKernel) is only needed before the first sweep of an iteration
        // and for the cells next to a connected face (their ghost neighbours are exchanged between
        // half sweeps).
        if (carry != nullptr) {
          double *o = carry + (planeBase + planeOf(q)) * (R::AN * kPCells) + cellG;
#pragma unroll
          for (int e = 0; e < neq; ++e) __stcg(o + e * kPCells, bsv[e]);
        }
      }
      NamedBarrier(1, C::threads);
    };

    // AITHER_B200_LUSGS_DBG=<file>: time line of every pencil (ticket drawn, planes 0, 32, 64, 96,
    // 128 reached, last plane done)
    auto stamp = [&](int slot) {
      if (dbg != nullptr && tid == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[ticket * 8 + slot] = t;
        if (slot == 0) {
          unsigned smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          dbg[ticket * 8 + 7] = bc.x + 1000 * bc.y + 1000000LL * smid;
        }
      }
    };
    stamp(0);
    NamedBarrier(1, C::threads);
    for (int q = 0; q < nSteps; ++q) {
      if ((q & 31) == 0 && q <= 128) stamp(1 + (q >> 5));
      step(q);
    }
    stamp(6);
    fills += nSteps;
  }
}

}  // namespace aither
