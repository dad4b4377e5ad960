// halo.cuh -- ghost-cell exchange across block connections (K12).
//
// Replaces the reference's slice -> MPI_Pack -> MPI_Sendrecv_replace -> InsertSlice path
// (ref: include/multiArray3d.hpp:790-926,1440-1550; src/boundaryConditions.cpp:833-858,
// 1016-1150,3006-3181; src/utility.cpp:400-423; src/gridLevel.cpp:297-312).
//
// Design (B200-first, not the reference's):
//   * every connection side is compiled ONCE, on the host, into two int32 index lists:
//       donor list     -- the donor block's slice (g interior layers, tangentially extended by the
//                         ghost layers, exactly connection::First/SecondSliceIndices) in the
//                         slice's own i-fastest order, as linear indices into the donor's fields;
//       acceptor list  -- (ghost cell, slice position) pairs: InsertSlice + GetSwapLoc with all 8
//                         orientations, lower/upper pairing and patch-border trimming resolved.
//     The device kernels are then pure gathers/scatters, component-major so every access is
//     coalesced along the list.
//   * the reference swaps connection after connection, each one "slice both sides, then insert
//     both sides"; tangential ghosts make later connections read what earlier ones wrote. That
//     order is kept bit-exactly by levelling: a connection goes one level after the last earlier
//     connection it has a RAW / WAW / WAR overlap with. One pack launch + one unpack launch per
//     LEVEL (3 levels for a Cartesian decomposition) instead of two per connection.
//   * same-GPU connections: the unpack reads the partner's pack buffer directly.
//     cross-GPU connections: the pack buffer travels with ncclSend / ncclRecv (one group per
//     level) over NVLink; NCCL is bound at run time with dlopen so the library has no link-time
//     dependency on it and single-GPU use needs no NCCL at all.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/aither_gpu.h"
#include "layout.cuh"

namespace aither {

enum HaloField { kHaloState = 0, kHaloUpdate = 1, kHaloTurb = 2, kHaloVelGrad = 3, kHaloWallDist = 4, kHaloNumFields };

inline std::string &HaloErrorRef() {
  static thread_local std::string e;
  return e;
}
inline std::string HaloError() { return HaloErrorRef(); }

// ---- NCCL, bound at run time -------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };  // layout of ncclUniqueId (nccl.h)
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
constexpr int kNcclSum = 0, kNcclMax = 2;

inline NcclApi *Nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    HaloErrorRef() = std::string("cannot load libnccl.so.2: ") + dlerror();
    return nullptr;
  }
#define BIND(field, sym)                                                    \
  *reinterpret_cast<void **>(&api.field) = dlsym(api.lib, sym);             \
  if (!api.field) {                                                         \
    HaloErrorRef() = std::string("libnccl lacks ") + sym;                   \
    api.lib = nullptr;                                                      \
    return nullptr;                                                         \
  }
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(GroupStart, "ncclGroupStart");
  BIND(GroupEnd, "ncclGroupEnd");
  BIND(Send, "ncclSend");
  BIND(Recv, "ncclRecv");
  BIND(AllReduce, "ncclAllReduce");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  return &api;
}

// ---- host-side compilation of one connection ---------------------------------------------------
struct Box {
  int lo[3], hi[3];  // half-open cell ranges in (i, j, k)
};
inline bool Overlap(const Box &a, const Box &b) {
  for (int d = 0; d < 3; ++d)
    if (a.hi[d] <= b.lo[d] || b.hi[d] <= a.lo[d]) return false;
  return true;
}

// the per-side fields of a connection, reorderable like connection::SwapOrder
struct ConnView {
  int boundary[2], d1s[2], d1e[2], d2s[2], d2e[2], cs[2];
  int border[8];
  int orientation;
};
inline ConnView ViewOf(const aither_conn &c) {
  ConnView v;
  for (int s = 0; s < 2; ++s) {
    v.boundary[s] = c.boundary[s];
    v.d1s[s] = c.d1Start[s];
    v.d1e[s] = c.d1End[s];
    v.d2s[s] = c.d2Start[s];
    v.d2e[s] = c.d2End[s];
    v.cs[s] = c.constSurf[s];
  }
  for (int q = 0; q < 8; ++q) v.border[q] = c.patchBorder[q];
  v.orientation = c.orientation;
  return v;
}
// ref: src/boundaryConditions.cpp:341-364
inline void SwapOrder(ConnView &v) {
  std::swap(v.boundary[0], v.boundary[1]);
  std::swap(v.d1s[0], v.d1s[1]);
  std::swap(v.d1e[0], v.d1e[1]);
  std::swap(v.d2s[0], v.d2s[1]);
  std::swap(v.d2e[0], v.d2e[1]);
  std::swap(v.cs[0], v.cs[1]);
  for (int q = 0; q < 4; ++q) std::swap(v.border[q], v.border[q + 4]);
  if (v.orientation == 4) v.orientation = 5;
  else if (v.orientation == 5) v.orientation = 4;
}
// surface type -> which of (i, j, k) are direction 3 (normal), 1 and 2
// (ref: src/boundaryConditions.cpp:875-968: i-surface: 1 = j, 2 = k; j: 1 = k, 2 = i; k: 1 = i, 2 = j)
inline void Dirs(int boundary, int *d3, int *d1, int *d2) {
  *d3 = (boundary - 1) / 2;
  *d1 = (*d3 + 1) % 3;
  *d2 = (*d3 + 2) % 3;
}
// the slice a side donates: connection::First/SecondSliceIndices (:1016-1150)
inline Box DonorBox(const ConnView &v, int s, int g) {
  int d3, d1, d2;
  Dirs(v.boundary[s], &d3, &d1, &d2);
  const int upLow = (v.boundary[s] % 2 == 0) ? -g : 0;
  Box b;
  b.lo[d3] = v.cs[s] + upLow;
  b.hi[d3] = b.lo[d3] + g;
  b.lo[d1] = v.d1s[s] - g;
  b.hi[d1] = v.d1e[s] + g;
  b.lo[d2] = v.d2s[s] - g;
  b.hi[d2] = v.d2e[s] + g;
  return b;
}

struct AcceptorMap {
  std::vector<int> cell;   // (i, j, k) triples of the ghost cells written
  std::vector<int> slice;  // position of the donor value inside the slice (i-fastest)
  Box written;
};

// InsertSlice (include/multiArray3d.hpp:876-926) with the connection adjusted by AdjustForSlice
// (src/boundaryConditions.cpp:833-858) and indices from GetSwapLoc (:3006-3181), for the side
// `acc` accepting the slice donated by side 1 - acc.
inline AcceptorMap BuildAcceptor(const aither_conn &c, int acc, int g) {
  ConnView v = ViewOf(c);
  const Box donor = DonorBox(v, 1 - acc, g);
  const int sn[3] = {donor.hi[0] - donor.lo[0], donor.hi[1] - donor.lo[1],
                     donor.hi[2] - donor.lo[2]};
  if (acc == 1) SwapOrder(v);  // the block inserted into is always "first"
  // AdjustForSlice
  const int blkStart = (v.boundary[0] % 2 == 0) ? v.cs[0] : -g;
  v.cs[1] = 0;
  v.cs[0] = blkStart;
  v.d1e[1] = v.d1e[1] - v.d1s[1] + 2 * g;
  v.d1e[0] += g;
  v.d1s[1] = 0;
  v.d1s[0] -= g;
  v.d2e[1] = v.d2e[1] - v.d2s[1] + 2 * g;
  v.d2e[0] += g;
  v.d2s[1] = 0;
  v.d2s[0] -= g;

  const int len1 = v.d1e[0] - v.d1s[0], len2 = v.d2e[0] - v.d2s[0];
  const int adjS1 = v.border[0] ? g : 0, adjE1 = v.border[1] ? g : 0;
  const int adjS2 = v.border[2] ? g : 0, adjE2 = v.border[3] ? g : 0;
  int a3, a1, a2, s3, s1, s2;
  Dirs(v.boundary[0], &a3, &a1, &a2);
  Dirs(v.boundary[1], &s3, &s1, &s2);
  const int o = v.orientation;
  const bool swap12 = o == 2 || o == 4 || o == 5 || o == 7;
  const bool sameSense = (v.boundary[0] + v.boundary[1]) % 2 == 0;  // lower/lower or upper/upper
  const int d3 = g;
  AcceptorMap m;
  for (int q = 0; q < 3; ++q) {
    m.written.lo[q] = 1 << 30;
    m.written.hi[q] = -(1 << 30);
  }
  for (int l3 = 0; l3 < d3; ++l3) {
    for (int l2 = adjS2; l2 < len2 - adjE2; ++l2) {
      for (int l1 = adjS1; l1 < len1 - adjE1; ++l1) {
        int A[3], S[3];
        // acceptor ("first"): IsLowerFirst() tests constSurf == 0, which after AdjustForSlice
        // holds only for an upper surface at 0 -- both branches reduce to cs + l3 (:3026)
        A[a1] = v.d1s[0] + l1;
        A[a2] = v.d2s[0] + l2;
        A[a3] = (v.cs[0] == 0) ? l3 - g : v.cs[0] + l3;
        // donor slice ("second"): slice has no ghosts, starts at 0
        if (swap12) {
          S[s2] = (o == 5 || o == 7) ? v.d2e[1] - 1 - l1 : v.d2s[1] + l1;
          S[s1] = (o == 4 || o == 7) ? v.d1e[1] - 1 - l2 : v.d1s[1] + l2;
        } else if (s3 == 0) {  // i-patch: 6/8 reverse direction 1, 3/8 direction 2 (:3065-3075)
          S[s1] = (o == 6 || o == 8) ? v.d1e[1] - 1 - l1 : v.d1s[1] + l1;
          S[s2] = (o == 3 || o == 8) ? v.d2e[1] - 1 - l2 : v.d2s[1] + l2;
        } else {  // j- and k-patches: 3/8 reverse direction 1, 6/8 direction 2 (:3108-3118,:3152)
          S[s1] = (o == 3 || o == 8) ? v.d1e[1] - 1 - l1 : v.d1s[1] + l1;
          S[s2] = (o == 6 || o == 8) ? v.d2e[1] - 1 - l2 : v.d2s[1] + l2;
        }
        S[s3] = sameSense ? d3 - l3 - 1 : l3;  // constSurf(second) == 0, slice ghosts == 0
        if (S[0] < 0 || S[0] >= sn[0] || S[1] < 0 || S[1] >= sn[1] || S[2] < 0 || S[2] >= sn[2]) {
          m.cell.clear();
          m.slice.clear();
          m.slice.push_back(-1);  // flags a geometry mismatch to the caller
          return m;
        }
        m.cell.push_back(A[0]);
        m.cell.push_back(A[1]);
        m.cell.push_back(A[2]);
        m.slice.push_back(S[0] + sn[0] * (S[1] + sn[1] * S[2]));
        for (int q = 0; q < 3; ++q) {
          m.written.lo[q] = std::min(m.written.lo[q], A[q]);
          m.written.hi[q] = std::max(m.written.hi[q], A[q] + 1);
        }
      }
    }
  }
  return m;
}

// ---- device side -------------------------------------------------------------------------------
struct HaloJob {
  const int *idx;    // pack: donor cell index per slice position; unpack: ghost cell index
  const int *pos;    // unpack only: slice position of each ghost cell
  double *buf;       // component-major staging: buf[e * sliceCells + n]
  int n;             // list length
  int sliceCells;    // cells in the slice (component stride of buf)
  int block;         // local block whose field is read / written
};

constexpr int kHaloMaxBlocks = 64;
struct HaloFields {
  double *base[kHaloMaxBlocks];
  long long fs[kHaloMaxBlocks];
};

static __global__ void __launch_bounds__(256)
    HaloPackKernel(const HaloJob *__restrict__ jobs, HaloFields f, int nc) {
  const HaloJob j = jobs[blockIdx.y];
  const double *__restrict__ src = f.base[j.block];
  const long long fs = f.fs[j.block];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < j.n; n += gridDim.x * blockDim.x) {
    const int c = __ldg(j.idx + n);
    for (int e = 0; e < nc; ++e) j.buf[static_cast<long long>(e) * j.sliceCells + n] = src[e * fs + c];
  }
}
static __global__ void __launch_bounds__(256)
    HaloUnpackKernel(const HaloJob *__restrict__ jobs, HaloFields f, int nc) {
  const HaloJob j = jobs[blockIdx.y];
  double *__restrict__ dst = f.base[j.block];
  const long long fs = f.fs[j.block];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < j.n; n += gridDim.x * blockDim.x) {
    const int c = __ldg(j.idx + n);
    const int p = __ldg(j.pos + n);
    for (int e = 0; e < nc; ++e) dst[e * fs + c] = j.buf[static_cast<long long>(e) * j.sliceCells + p];
  }
}

struct HaloXfer {     // one direction of one connection: donor side -> acceptor side
  int conn = 0, acc = 0;
  int donorRank = 0, accRank = 0, donorBlock = -1, accBlock = -1;  // local block ids (-1: remote)
  int sliceCells = 0, nAcc = 0;
  int *dDonorIdx = nullptr, *dAccIdx = nullptr, *dAccPos = nullptr;
  double *dBuf = nullptr;  // pack target (donor local) or receive target (donor remote)
};
struct HaloLevel {
  std::vector<int> xfers;       // indices into HaloPlan::xfers, connection order
  HaloJob *dPack = nullptr, *dUnpack = nullptr;
  int nPack = 0, nUnpack = 0, maxPack = 0, maxUnpack = 0;
  bool anyRemote = false;
  // direct exchange over peer memory: job lists per buffer parity, the ranks this level writes
  // to / is written by
  HaloJob *dPackP2P[2] = {nullptr, nullptr}, *dUnpackP2P[2] = {nullptr, nullptr};
  std::vector<int> sendTo, recvFrom;
};
struct HaloPlan {
  int nConn = 0;
  int rank = 0, nRanks = 1;
  int maxComp = 0;
  void *comm = nullptr;
  std::vector<HaloXfer> xfers;
  std::vector<HaloLevel> levels;
  std::vector<void *> owned;    // device allocations
  long long bytesPerExchangeRemote = 0;  // doubles sent per component per exchange, for reports
  // what every rank knows about every connection (the plan is compiled from the global list):
  // slice cells of acceptor side `acc` of connection c at [2 c + acc], its level, the two ranks
  std::vector<long long> allSlice;
  std::vector<int> allLevel, allRank;
  // direct exchange over peer memory (HaloP2PEnable)
  bool p2p = false;
  int planId = 0;
  unsigned seq = 0;  // exchanges done on this plan: buffer parity and flag value
  unsigned *myFlags = nullptr;
  std::vector<unsigned *> peerFlags;
};

inline int HaloFail(const std::string &m) {
  HaloErrorRef() = m;
  return 1;
}

template <typename T>
inline int HaloUpload(HaloPlan &plan, const std::vector<T> &v, T **out) {
  *out = nullptr;
  if (v.empty()) return 0;
  if (cudaMalloc(out, sizeof(T) * v.size()) != cudaSuccess) return HaloFail("halo: cudaMalloc failed");
  plan.owned.push_back(*out);
  if (cudaMemcpy(*out, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    return HaloFail("halo: cudaMemcpy failed");
  return 0;
}

// Compile the connections this rank takes part in. `conns` is the full (global) list in the
// reference's order; every rank derives the same levels from it.
// `faceOnly`: keep only the ghost cells straight behind each patch (no tangential extension into
// the edge ghost cells). Those are all that the inviscid stencils, the implicit off-diagonals and
// the turbulence / gradient exchanges read; without the extension no two connections touch the
// same cell, every connection lands in ONE level (one pack + one NCCL group + one unpack per
// exchange instead of three for a Cartesian decomposition), and what is written there is
// bit-identical to the full plan's.
inline int HaloBuild(HaloPlan &plan, const std::vector<aither_conn> &conns,
                     const std::vector<const BlockDev *> &devs, const std::vector<int> &globalPos,
                     int maxComp, int g, int rank, int nRanks, void *ncclComm,
                     bool faceOnly = false) {
  (void)globalPos;
  plan.nConn = static_cast<int>(conns.size());
  plan.rank = rank;
  plan.nRanks = nRanks;
  plan.comm = ncclComm;
  plan.maxComp = maxComp;
  if (conns.empty()) return 0;
  if (static_cast<int>(devs.size()) > kHaloMaxBlocks)
    return HaloFail("halo: more than " + std::to_string(kHaloMaxBlocks) + " local blocks per GPU");
  const int nc = plan.nConn;
  // levels from read/write overlaps, over the global list
  struct RW { int rk[2], lb[2]; Box rd[2], wr[2]; };
  std::vector<RW> rw(nc);
  std::vector<AcceptorMap> maps(2 * static_cast<size_t>(nc));
  std::vector<Box> faceRead(2 * static_cast<size_t>(nc));
  for (int c = 0; c < nc; ++c) {
    const ConnView v = ViewOf(conns[c]);
    if (conns[c].orientation < 1 || conns[c].orientation > 8)
      return HaloFail("halo: connection " + std::to_string(c) + " has orientation outside 1..8");
    for (int s = 0; s < 2; ++s) {
      rw[c].rk[s] = conns[c].rank[s];
      rw[c].lb[s] = conns[c].localBlock[s];
      rw[c].rd[s] = DonorBox(v, s, g);
      maps[2 * c + s] = BuildAcceptor(conns[c], s, g);
      if (maps[2 * c + s].slice.size() == 1 && maps[2 * c + s].slice[0] < 0)
        return HaloFail("halo: connection " + std::to_string(c) + " patches do not match in size");
      rw[c].wr[s] = maps[2 * c + s].written;
      if (faceOnly) {
        // drop the pairs whose target lies outside the patch's own tangential range, then take
        // the bounding boxes of what is still written and of the donor cells still read
        AcceptorMap &m = maps[2 * c + s];
        int d3, d1, d2;
        Dirs(v.boundary[s], &d3, &d1, &d2);
        const Box don = DonorBox(v, 1 - s, g);
        const int sn0 = don.hi[0] - don.lo[0], sn1 = don.hi[1] - don.lo[1];
        AcceptorMap f;
        Box wr = {{1 << 30, 1 << 30, 1 << 30}, {-(1 << 30), -(1 << 30), -(1 << 30)}};
        Box rd = wr;
        std::vector<int> donorCell;
        for (size_t n = 0; n < m.slice.size(); ++n) {
          const int *cc = &m.cell[3 * n];
          if (cc[d1] < v.d1s[s] || cc[d1] >= v.d1e[s] || cc[d2] < v.d2s[s] || cc[d2] >= v.d2e[s])
            continue;
          f.cell.insert(f.cell.end(), cc, cc + 3);
          const int sp = m.slice[n];
          const int dc[3] = {don.lo[0] + sp % sn0, don.lo[1] + (sp / sn0) % sn1,
                             don.lo[2] + sp / (sn0 * sn1)};
          donorCell.insert(donorCell.end(), dc, dc + 3);
          for (int q = 0; q < 3; ++q) {
            wr.lo[q] = std::min(wr.lo[q], cc[q]);
            wr.hi[q] = std::max(wr.hi[q], cc[q] + 1);
            rd.lo[q] = std::min(rd.lo[q], dc[q]);
            rd.hi[q] = std::max(rd.hi[q], dc[q] + 1);
          }
        }
        // slice positions inside the clipped donor box (the slice that is packed and sent)
        const int rn0 = rd.hi[0] - rd.lo[0], rn1 = rd.hi[1] - rd.lo[1];
        for (size_t n = 0; n < donorCell.size() / 3; ++n) {
          const int *dc = &donorCell[3 * n];
          f.slice.push_back((dc[0] - rd.lo[0]) + rn0 * ((dc[1] - rd.lo[1]) + rn1 * (dc[2] - rd.lo[2])));
        }
        f.written = wr;
        m = f;
        rw[c].wr[s] = wr;
        faceRead[2 * c + (1 - s)] = rd;  // what the donor side (1 - s) must provide
      }
    }
  }
  if (faceOnly)
    for (int c = 0; c < nc; ++c)
      for (int s = 0; s < 2; ++s) rw[c].rd[s] = faceRead[2 * c + s];
  std::vector<int> level(nc, 0);
  int nLevels = 0;
  for (int c = 0; c < nc; ++c) {
    for (int p = 0; p < c; ++p) {
      bool conflict = false;
      for (int s = 0; s < 2 && !conflict; ++s)
        for (int t = 0; t < 2 && !conflict; ++t) {
          if (rw[c].rk[s] != rw[p].rk[t] || rw[c].lb[s] != rw[p].lb[t]) continue;
          conflict = Overlap(rw[p].wr[t], rw[c].rd[s]) || Overlap(rw[p].wr[t], rw[c].wr[s]) ||
                     Overlap(rw[p].rd[t], rw[c].wr[s]);
        }
      if (conflict) level[c] = std::max(level[c], level[p] + 1);
    }
    nLevels = std::max(nLevels, level[c] + 1);
  }
  plan.levels.resize(nLevels);
  plan.allSlice.assign(2 * static_cast<size_t>(nc), 0);
  plan.allLevel = level;
  plan.allRank.assign(2 * static_cast<size_t>(nc), 0);
  for (int c = 0; c < nc; ++c)
    for (int acc = 0; acc < 2; ++acc) {
      const Box &db = rw[c].rd[1 - acc];
      plan.allSlice[2 * c + acc] = static_cast<long long>(db.hi[0] - db.lo[0]) * (db.hi[1] - db.lo[1]) *
                                   (db.hi[2] - db.lo[2]);
      plan.allRank[2 * c + acc] = conns[c].rank[acc];
    }

  bool needNccl = false;
  for (int c = 0; c < nc; ++c) {
    const aither_conn &cn = conns[c];
    if (cn.rank[0] != rank && cn.rank[1] != rank) continue;
    // side 0 accepts first, then side 1 (SwapSliceLocal puts into array1 first, :821-822)
    for (int acc = 0; acc < 2; ++acc) {
      const int don = 1 - acc;
      HaloXfer x;
      x.conn = c;
      x.acc = acc;
      x.donorRank = cn.rank[don];
      x.accRank = cn.rank[acc];
      x.donorBlock = cn.rank[don] == rank ? cn.localBlock[don] : -1;
      x.accBlock = cn.rank[acc] == rank ? cn.localBlock[acc] : -1;
      if (x.donorBlock >= static_cast<int>(devs.size()) || x.accBlock >= static_cast<int>(devs.size()))
        return HaloFail("halo: connection " + std::to_string(c) + " names a local block this rank does not own");
      const Box &db = rw[c].rd[don];
      const int sn[3] = {db.hi[0] - db.lo[0], db.hi[1] - db.lo[1], db.hi[2] - db.lo[2]};
      x.sliceCells = sn[0] * sn[1] * sn[2];
      if (x.donorBlock >= 0) {
        const BlockDev &b = *devs[x.donorBlock];
        if (b.fs >= (1LL << 31)) return HaloFail("halo: block too large for 32-bit halo indices");
        if (db.lo[0] < -b.g || db.hi[0] > b.ni + b.g || db.lo[1] < -b.g || db.hi[1] > b.nj + b.g ||
            db.lo[2] < -b.g || db.hi[2] > b.nk + b.g)
          return HaloFail("halo: connection " + std::to_string(c) + " slice leaves its block");
        std::vector<int> idx(x.sliceCells);
        size_t n = 0;
        for (int k = db.lo[2]; k < db.hi[2]; ++k)
          for (int j = db.lo[1]; j < db.hi[1]; ++j)
            for (int i = db.lo[0]; i < db.hi[0]; ++i) idx[n++] = static_cast<int>(CellIdx(b, i, j, k));
        if (HaloUpload(plan, idx, &x.dDonorIdx)) return 1;
      }
      if (x.accBlock >= 0) {
        const BlockDev &b = *devs[x.accBlock];
        if (b.fs >= (1LL << 31)) return HaloFail("halo: block too large for 32-bit halo indices");
        const AcceptorMap &m = maps[2 * c + acc];
        x.nAcc = static_cast<int>(m.slice.size());
        std::vector<int> idx(x.nAcc);
        for (int n = 0; n < x.nAcc; ++n) {
          const int i = m.cell[3 * n], j = m.cell[3 * n + 1], k = m.cell[3 * n + 2];
          if (i < -b.g || i >= b.ni + b.g || j < -b.g || j >= b.nj + b.g || k < -b.g || k >= b.nk + b.g)
            return HaloFail("halo: connection " + std::to_string(c) + " writes outside its block");
          idx[n] = static_cast<int>(CellIdx(b, i, j, k));
        }
        if (HaloUpload(plan, idx, &x.dAccIdx)) return 1;
        if (HaloUpload(plan, m.slice, &x.dAccPos)) return 1;
      }
      if (cudaMalloc(&x.dBuf, sizeof(double) * static_cast<size_t>(x.sliceCells) * maxComp) != cudaSuccess)
        return HaloFail("halo: cudaMalloc failed");
      plan.owned.push_back(x.dBuf);
      if (x.donorBlock < 0 || x.accBlock < 0) {
        needNccl = true;
        plan.levels[level[c]].anyRemote = true;
        if (x.donorBlock >= 0) plan.bytesPerExchangeRemote += x.sliceCells;
      }
      plan.levels[level[c]].xfers.push_back(static_cast<int>(plan.xfers.size()));
      plan.xfers.push_back(x);
    }
  }
  if (needNccl) {
    if (!ncclComm) return HaloFail("halo: connections cross ranks but no NCCL communicator was given");
    if (!Nccl()) return 1;
  }
  for (auto &lv : plan.levels) {
    std::vector<HaloJob> pack, unpack;
    for (int xi : lv.xfers) {
      const HaloXfer &x = plan.xfers[xi];
      if (x.donorBlock >= 0) {
        pack.push_back({x.dDonorIdx, nullptr, x.dBuf, x.sliceCells, x.sliceCells, x.donorBlock});
        lv.maxPack = std::max(lv.maxPack, x.sliceCells);
      }
      if (x.accBlock >= 0) {
        unpack.push_back({x.dAccIdx, x.dAccPos, x.dBuf, x.nAcc, x.sliceCells, x.accBlock});
        lv.maxUnpack = std::max(lv.maxUnpack, x.nAcc);
      }
    }
    lv.nPack = static_cast<int>(pack.size());
    lv.nUnpack = static_cast<int>(unpack.size());
    if (HaloUpload(plan, pack, &lv.dPack)) return 1;
    if (HaloUpload(plan, unpack, &lv.dUnpack)) return 1;
  }
  return 0;
}

// ---- direct exchange over peer memory ---------------------------------------------------------
// ncclSend / ncclRecv cost ~95 us per exchange on NVLink whatever the size (six exchanges per
// iteration: 0.57 ms of 8.5 at 8 GPUs). Here the pack kernel of the donor rank writes the slice
// straight into the acceptor rank's receive buffer (peer-mapped over NVLink: cudaIpc*), a stream
// memory operation raises a flag in the acceptor's memory behind it (cuStreamWriteValue32 fences
// the writes issued before it) and the acceptor's stream waits for the flag before it unpacks
// (cuStreamWaitValue32): no kernel spins, no proxy thread. Every rank's receive buffers and flags
// live in ONE allocation (one IPC handle per rank) whose layout every rank computes from the
// global connection list. Two buffers per slice, used alternately: a rank cannot start exchange
// n + 2 of a plan before its own exchange n + 1 has seen the partner's flag, which the partner
// raises after unpacking exchange n.
constexpr int kP2PMaxLevels = 8;
constexpr size_t kP2PFlagBytes = 64 * 1024;
constexpr int kP2PHandleBytes = 64;  // sizeof(cudaIpcMemHandle_t)

inline size_t P2PSliceBytes(const HaloPlan &plan, int c, int acc) {
  const size_t b = sizeof(double) * static_cast<size_t>(plan.allSlice[2 * c + acc]) * plan.maxComp;
  return (b + 255) / 256 * 256;
}
// offset of the receive buffers (parity 0, then parity 1) of acceptor side `acc` of connection
// `c` of plan `pi` in the arena of rank R; (pi, c, acc) = (-1, ., .): the arena's size
inline size_t P2POffset(HaloPlan *const *plans, int nPlans, int R, int pi, int c, int acc) {
  size_t off = kP2PFlagBytes;
  for (int q = 0; q < nPlans; ++q) {
    const HaloPlan &pl = *plans[q];
    for (int cc = 0; cc < pl.nConn; ++cc)
      for (int a = 0; a < 2; ++a) {
        if (pl.allRank[2 * cc + a] != R || pl.allRank[2 * cc + (1 - a)] == R) continue;
        if (q == pi && cc == c && a == acc) return off;
        off += 2 * P2PSliceBytes(pl, cc, a);
      }
  }
  return off;
}
inline size_t P2PFlagSlot(const HaloPlan &plan, int level, int srcRank) {
  return (static_cast<size_t>(plan.planId) * kP2PMaxLevels + level) * plan.nRanks + srcRank;
}

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
inline StreamValue32Fn StreamValueFn(const char *name) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    return reinterpret_cast<StreamValue32Fn>(p);
  return nullptr;
}

// switch the plans to the direct exchange: arena[r] = rank r's arena as this process sees it
inline int HaloP2PEnable(HaloPlan *const *plans, int nPlans, const std::vector<unsigned char *> &arena) {
  for (int pi = 0; pi < nPlans; ++pi) {
    HaloPlan &plan = *plans[pi];
    if (plan.nConn == 0) continue;
    if (static_cast<int>(plan.levels.size()) > kP2PMaxLevels)
      return HaloFail("halo: more exchange levels than the direct exchange keeps flags for");
    if (kP2PMaxLevels * static_cast<size_t>(nPlans) * plan.nRanks * sizeof(unsigned) > kP2PFlagBytes)
      return HaloFail("halo: too many ranks for the direct exchange's flag block");
    plan.planId = pi;
    plan.myFlags = reinterpret_cast<unsigned *>(arena[plan.rank]);
    plan.peerFlags.resize(plan.nRanks);
    for (int r = 0; r < plan.nRanks; ++r) plan.peerFlags[r] = reinterpret_cast<unsigned *>(arena[r]);
    for (auto &lv : plan.levels) {
      lv.sendTo.clear();
      lv.recvFrom.clear();
      for (int par = 0; par < 2; ++par) {
        std::vector<HaloJob> pack, unpack;
        for (int xi : lv.xfers) {
          const HaloXfer &x = plan.xfers[xi];
          const size_t sb = P2PSliceBytes(plan, x.conn, x.acc);
          if (x.donorBlock >= 0) {
            double *buf = x.dBuf;
            if (x.accBlock < 0) {
              buf = reinterpret_cast<double *>(arena[x.accRank] +
                                               P2POffset(plans, nPlans, x.accRank, pi, x.conn, x.acc) +
                                               par * sb);
              if (par == 0 && std::find(lv.sendTo.begin(), lv.sendTo.end(), x.accRank) == lv.sendTo.end())
                lv.sendTo.push_back(x.accRank);
            }
            pack.push_back({x.dDonorIdx, nullptr, buf, x.sliceCells, x.sliceCells, x.donorBlock});
          }
          if (x.accBlock >= 0) {
            double *buf = x.dBuf;
            if (x.donorBlock < 0) {
              buf = reinterpret_cast<double *>(arena[plan.rank] +
                                               P2POffset(plans, nPlans, plan.rank, pi, x.conn, x.acc) +
                                               par * sb);
              if (par == 0 && std::find(lv.recvFrom.begin(), lv.recvFrom.end(), x.donorRank) == lv.recvFrom.end())
                lv.recvFrom.push_back(x.donorRank);
            }
            unpack.push_back({x.dAccIdx, x.dAccPos, buf, x.nAcc, x.sliceCells, x.accBlock});
          }
        }
        if (HaloUpload(plan, pack, &lv.dPackP2P[par])) return 1;
        if (HaloUpload(plan, unpack, &lv.dUnpackP2P[par])) return 1;
      }
    }
    plan.seq = 0;
    plan.p2p = true;
  }
  return 0;
}

// Exchange `nc` components of one field of every local block (base[b] = component 0 of block b).
// Asynchronous on `stream`. `launches` counts kernels launched.
inline int HaloExchange(HaloPlan &plan, const HaloFields &f, int nc, cudaStream_t stream,
                        long long *launches, long long *packLaunches) {
  if (plan.nConn == 0) return 0;
  if (nc > plan.maxComp) return HaloFail("halo: field has more components than the plan's buffers");
  if (plan.p2p) {
    static StreamValue32Fn writeFn = StreamValueFn("cuStreamWriteValue32");
    static StreamValue32Fn waitFn = StreamValueFn("cuStreamWaitValue32");
    if (!writeFn || !waitFn) return HaloFail("halo: stream memory operations are not available");
    const unsigned seq = ++plan.seq;
    const int par = static_cast<int>(seq & 1u);
    for (size_t L = 0; L < plan.levels.size(); ++L) {
      HaloLevel &lv = plan.levels[L];
      if (lv.nPack > 0) {
        const dim3 grid(std::min((lv.maxPack + 255) / 256, 148 * 4), lv.nPack);
        HaloPackKernel<<<grid, 256, 0, stream>>>(lv.dPackP2P[par], f, nc);
        if (launches) ++*launches;
        if (packLaunches) ++*packLaunches;
      }
      for (int r : lv.sendTo)
        if (writeFn(stream, reinterpret_cast<CUdeviceptr>(plan.peerFlags[r] + P2PFlagSlot(plan, static_cast<int>(L), plan.rank)),
                    seq, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
          return HaloFail("halo: cuStreamWriteValue32 on a peer's flag failed");
      for (int r : lv.recvFrom)
        if (waitFn(stream, reinterpret_cast<CUdeviceptr>(plan.myFlags + P2PFlagSlot(plan, static_cast<int>(L), r)),
                   seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
          return HaloFail("halo: cuStreamWaitValue32 failed");
      if (lv.nUnpack > 0) {
        const dim3 grid(std::min((lv.maxUnpack + 255) / 256, 148 * 4), lv.nUnpack);
        HaloUnpackKernel<<<grid, 256, 0, stream>>>(lv.dUnpackP2P[par], f, nc);
        if (launches) ++*launches;
        if (packLaunches) ++*packLaunches;
      }
    }
    if (cudaGetLastError() != cudaSuccess) return HaloFail("halo: kernel launch failed");
    return 0;
  }
  NcclApi *api = nullptr;
  for (auto &lv : plan.levels) {
    if (lv.nPack > 0) {
      const dim3 grid(std::min((lv.maxPack + 255) / 256, 148 * 4), lv.nPack);
      HaloPackKernel<<<grid, 256, 0, stream>>>(lv.dPack, f, nc);
      if (launches) ++*launches;
      if (packLaunches) ++*packLaunches;
    }
    if (lv.anyRemote) {
      if (!api) api = Nccl();
      if (!api) return 1;
      int rc = api->GroupStart();
      for (int xi : lv.xfers) {
        const HaloXfer &x = plan.xfers[xi];
        const size_t cnt = static_cast<size_t>(x.sliceCells) * nc;
        if (x.donorBlock >= 0 && x.accBlock < 0 && rc == 0)
          rc = api->Send(x.dBuf, cnt, kNcclFloat64, x.accRank, plan.comm, stream);
        if (x.donorBlock < 0 && x.accBlock >= 0 && rc == 0)
          rc = api->Recv(x.dBuf, cnt, kNcclFloat64, x.donorRank, plan.comm, stream);
      }
      const int rc2 = api->GroupEnd();
      if (rc != 0 || rc2 != 0)
        return HaloFail(std::string("halo: NCCL send/recv failed: ") +
                        api->GetErrorString(rc != 0 ? rc : rc2));
    }
    if (lv.nUnpack > 0) {
      const dim3 grid(std::min((lv.maxUnpack + 255) / 256, 148 * 4), lv.nUnpack);
      HaloUnpackKernel<<<grid, 256, 0, stream>>>(lv.dUnpack, f, nc);
      if (launches) ++*launches;
      if (packLaunches) ++*packLaunches;
    }
  }
  if (cudaGetLastError() != cudaSuccess) return HaloFail("halo: kernel launch failed");
  return 0;
}

inline void HaloDestroy(HaloPlan &plan) {
  for (void *p : plan.owned) cudaFree(p);
  plan.owned.clear();
  plan.xfers.clear();
  plan.levels.clear();
  plan.nConn = 0;
}

}  // namespace aither
