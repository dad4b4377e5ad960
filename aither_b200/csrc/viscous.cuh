// viscous.cuh -- laminar Navier-Stokes terms of the residual (K2/K3) and what they need around
// them: Sutherland transport, viscous-wall and edge ghost cells, viscous spectral radii.
//
// Reference path (mnucci32/aither v0.10.0): procBlock::CalcResidualNoSource viscous branch
// (src/procBlock.cpp:6125-6137) = AssignViscousGhostCells (:2760-3025) -> UpdateAuxillaryVariables
// (:6171-6190) -> CalcViscFluxI/J/K (:1233-2200) with CalcGradsI/J/K (:5173-5780),
// viscousFlux::CalcFlux (src/viscousFlux.cpp:58-135), ViscCellSpectralRadius
// (include/spectralRadius.hpp:94-124).
//
// Device design: a face's viscous flux is computed ONCE, by the thread of the cell above it
// (ViscFaceKernel, all three directions in one pass over the block, 10-cell Green-Gauss stencil
// served by L1/L2), and parked in scratch fields that are idle during the residual phase (the
// implicit update's ping-pong buffer and the matrix-residual field); ViscAccumKernel then adds
// the six face fluxes of every cell to its residual in the reference's order (+lower, -upper for
// i, j, k) together with the viscous spectral radius and its share of the scalar
// diagonal. Owner-writes, no atomics, run-to-run identical.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "walllaw.cuh"

namespace aither {

// viscous-wall ghost state (ref: src/ghostStates.cpp:134-281): velocity mirrored about the wall
// velocity; adiabatic keeps rho and p, isothermal / heat-flux walls set the ghost temperature and
// take rho from p = rho R T. Low-Re treatment: k = 0 and the omega wall value; with the wall law
// (bc.isWallLaw, `nA` = unit normal out of the domain) the ghost temperature, k and omega come
// from the law unless its y+ fell below 10 (wallVars::SwitchToLowRe). `wvOut`: the wall variables
// of this face, y+ = 0 when the wall has no wall law.
template <int NS, int NT>
AITHER_HD void ViscousWallGhost(const Gas &g, const Transport &tr, const double *interior,
                                const aither_bc_state &bc, double wallDist, double *ghost,
                                double nuW = 0.0, int layer = 1, const double *nA = nullptr,
                                bool isLower = false, WallVars *wvOut = nullptr) {
  using E = Eq<NS, NT>;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) ghost[e] = interior[e];
  ghost[E::imx] = 2.0 * bc.velocity[0] - interior[E::imx];
  ghost[E::imy] = 2.0 * bc.velocity[1] - interior[E::imy];
  ghost[E::imz] = 2.0 * bc.velocity[2] - interior[E::imz];
  WallVars wv;
  wv.yplus = 0.0;
  const bool wallLaw = bc.isWallLaw && nA != nullptr;
  bool lowRe = true;
  if (wallLaw) {
    WallLawEval<NS, NT>(g, tr, bc,
                        bc.isIsothermal ? kWallIsothermal
                                        : (bc.isConstantHeatFlux ? kWallHeatFlux : kWallAdiabatic),
                        interior, wallDist, nA, isLower, wv);
    lowRe = wv.SwitchToLowRe();
  }
  if (bc.isIsothermal || bc.isConstantHeatFlux) {
    const double tInt = Temperature<NS>(g, interior);
    double tGhost;
    if (bc.isIsothermal) {
      if (!lowRe) {  // wall-law heat flux through laminar + turbulent conductivity (:160-171)
        const double kappa = MixtureEffConductivity<NS>(tr, wv.t, interior) +
                             wv.mut * Mixture<NS>(g, interior).cp / TurbPrandtl(tr.turbModel);
        tGhost = bc.temperature - wv.heatFlux / kappa * 2.0 * wallDist;
      } else {
        tGhost = 2.0 * bc.temperature - tInt;
      }
    } else {
      if (!lowRe) {  // wall-law wall temperature (:212-219)
        tGhost = 2.0 * wv.t - tInt;
      } else {
        const double kappa = MixtureEffConductivity<NS>(tr, tInt, interior);
        tGhost = tInt - bc.heatFlux / kappa * 2.0 * wallDist;
      }
    }
    const double rhoInt = SpeciesSum<NS>(interior);
    double R = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) R += interior[q] / rhoInt * g.R[q];
    const double rho = ghost[E::ie] / (R * tGhost);
#pragma unroll
    for (int q = 0; q < NS; ++q) ghost[q] = rho * (interior[q] / rhoInt);
  }
  if (NT > 1) {
    constexpr int iw = E::it + (NT > 1 ? 1 : 0);
    if (!lowRe) {  // k and omega of the law at the wall (:173-180, :221-228, :248-255)
      ghost[E::it] = 2.0 * wv.tke - interior[E::it];
      ghost[iw] = 2.0 * wv.sdr - interior[iw];
      if (layer > 1) {
        ghost[E::it] = layer * ghost[E::it] - wv.tke;
        ghost[iw] = layer * ghost[iw] - wv.sdr;
      }
    } else {
      // k = 0 at the wall; omega_wall = scaling^2 60 nu_w / (beta d^2) (ref: :262-281)
      ghost[E::it] = -1.0 * interior[E::it];
      const double wWall = tr.scaling * tr.scaling * 60.0 * nuW /
                           (wallDist * wallDist * TurbWallBeta(tr.turbModel));
      ghost[iw] = 2.0 * wWall - interior[iw];
      if (layer > 1) ghost[iw] = layer * ghost[iw] - wWall;
    }
  }
  if (wvOut) *wvOut = wv;
}

// ---- K11b: viscous-wall ghost cells (ref: src/procBlock.cpp:2760-2833) -------------------------
template <int NS, int NT>
__global__ void ViscousWallKernel(BlockDev b, Params p, const SurfDev *__restrict__ surfs, int nsurf,
                                  const aither_bc_state *__restrict__ bcs, long long total) {
  using E = Eq<NS, NT>;
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t >= total) return;
  int s = 0;
  while (s + 1 < nsurf && surfs[s + 1].faceOffset <= t) ++s;
  const SurfDev sf = surfs[s];
  if (sf.type != AITHER_BC_VISCOUS_WALL) return;
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
  int c[3];
  c[d1] = sf.lo[d1] + a1;
  c[d2] = sf.lo[d2] + a2;
  c[d3] = iCell;
  double interior[E::neq], ghost[E::neq];
  LoadCell<E::neq>(b.state, b.fs, CellIdx(b, c[0], c[1], c[2]), interior);
  c[d3] = aCell;
  const long long aidx = CellIdx(b, c[0], c[1], c[2]);
  const double wd = b.wallDist ? b.wallDist[aidx] : 0.0;
  // nu of the wall-adjacent cell from the STORED viscosity, i.e. the previous evaluation's
  // (AssignViscousGhostCells runs before UpdateAuxillaryVariables; ref src/procBlock.cpp:2814-2822)
  double nuW = 0.0;
  if (NT > 0) {
    double rhoA = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) rhoA += __ldg(b.state + q * b.fs + aidx);
    nuW = b.viscosity[aidx] / rhoA;
  }
  const aither_bc_state &bc = bcs[sf.bcIndex];
  if (bc.isWallLaw) {
    // unit normal out of the domain and the record of this face (first layer only;
    // ref: src/procBlock.cpp:6287-6290)
    c[d3] = r3;
    const long long fidx = CellIdx(b, c[0], c[1], c[2]);
    const bool isLower = sf.surfType % 2 == 1;
    double nA[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double a = __ldg(b.fA[d3] + q * b.fs + fidx);
      nA[q] = isLower ? -1.0 * a : a;
    }
    WallVars wv;
    ViscousWallGhost<NS, NT>(p.gas, p.tr, interior, bc, wd, ghost, nuW, layer, nA, isLower, &wv);
    if (layer == 1 && b.wallVars) {
      double *w = b.wallVars + kWallVarsStride * (sf.faceOffset / b.g + a1 + static_cast<long long>(n1) * a2);
      w[kWvYplus] = wv.yplus;
      w[kWvTau] = wv.tau[0];
      w[kWvTau + 1] = wv.tau[1];
      w[kWvTau + 2] = wv.tau[2];
      w[kWvHeatFlux] = wv.heatFlux;
      w[kWvMu] = wv.mu;
      w[kWvMut] = wv.mut;
      w[kWvRho] = wv.rho;
      w[kWvT] = wv.t;
      w[kWvTke] = wv.tke;
      w[kWvSdr] = wv.sdr;
      w[kWvVelWall] = bc.velocity[0];
      w[kWvVelWall + 1] = bc.velocity[1];
      w[kWvVelWall + 2] = bc.velocity[2];
      w[kWvUtau] = wv.utau;
    }
  } else {
    ViscousWallGhost<NS, NT>(p.gas, p.tr, interior, bc, wd, ghost, nuW, layer);
  }
  c[d3] = gCell;
  StoreCell<E::neq>(b.state, b.fs, CellIdx(b, c[0], c[1], c[2]), ghost);
}

// ---- K11c: edge ghost cells --------------------------------------------------------------------
// ref: src/procBlock.cpp:2565-2703 (AssignInviscidGhostCellsEdge; VISCOUS = false) and
// :2873-3025 (AssignViscousGhostCellsEdge; VISCOUS = true). One thread per (edge direction, one of
// its 4 edges, position along the edge); the g x g cells of that position are filled in the
// reference's layer order because later layers read earlier ones.
struct EdgeSurf {
  int type, surfType, tag, bcIndex;
  int faceBase;      // first boundary-face record of the surface in BlockDev::wallVars (-1: connection)
  int lo[3], hi[3];  // node-index ranges of the surface as given (imin..kmax)
};
__device__ __forceinline__ void DirIjk(int dd, int d1, int d2, int d3, int *c) {
  // multiArray3d::operator()(dir, d1, d2, d3): include/multiArray3d.hpp:231-243
  if (dd == 0) { c[0] = d1; c[1] = d2; c[2] = d3; }
  else if (dd == 1) { c[0] = d3; c[1] = d1; c[2] = d2; }
  else { c[0] = d2; c[1] = d3; c[2] = d1; }
}
__device__ __forceinline__ int FindSurface(const EdgeSurf *surfs, int nsurf, const int *c, int surf) {
  // ref: src/boundaryConditions.cpp:109-185 (GetBCSurface)
  const int sd = (surf - 1) / 2;
  for (int s = 0; s < nsurf; ++s) {
    if ((surfs[s].surfType - 1) / 2 != sd) continue;
    bool in = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (q == sd) in = in && c[q] >= surfs[s].lo[q] && c[q] <= surfs[s].hi[q];
      else in = in && c[q] >= surfs[s].lo[q] && c[q] < surfs[s].hi[q];
    }
    if (in) return s;
  }
  return -1;
}
template <int NS, int NT, bool VISCOUS>
__global__ void EdgeKernel(BlockDev b, Params p, const EdgeSurf *__restrict__ surfs, int nsurf,
                           const aither_bc_state *__restrict__ bcs) {
  using E = Eq<NS, NT>;
  const int nd[3] = {b.ni, b.nj, b.nk};
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int dd = 0;
  for (; dd < 3; ++dd) {
    if (t < 4 * nd[dd]) break;
    t -= 4 * nd[dd];
  }
  if (dd == 3) return;
  const int cc = t / nd[dd], d1 = t % nd[dd];
  const int max2 = nd[(dd + 1) % 3], max3 = nd[(dd + 2) % 3];
  const int surfStart2 = 2 * ((dd + 1) % 3) + 1, surfStart3 = 2 * ((dd + 2) % 3) + 1;
  const int fd2 = (dd + 1) % 3, fd3 = (dd + 2) % 3;
  const bool upper2 = cc > 1, upper3 = cc % 2 == 1;
  const int surf2 = upper2 ? surfStart2 + 1 : surfStart2;
  const int surf3 = upper3 ? surfStart3 + 1 : surfStart3;
  const int cFaceD2_2 = upper2 ? max2 : 0, cFaceD2_3 = upper3 ? max3 - 1 : 0;
  const int cFaceD3_2 = upper2 ? max2 - 1 : 0, cFaceD3_3 = upper3 ? max3 : 0;
  int c2[3], c3[3];
  DirIjk(dd, d1, cFaceD2_2, cFaceD2_3, c2);
  DirIjk(dd, d1, cFaceD3_2, cFaceD3_3, c3);
  const int s2 = FindSurface(surfs, nsurf, c2, surf2), s3 = FindSurface(surfs, nsurf, c3, surf3);
  if (s2 < 0 || s3 < 0) return;
  int bc2 = surfs[s2].type, bc3 = surfs[s3].type;
  if (!VISCOUS) {
    if (bc2 == AITHER_BC_VISCOUS_WALL) bc2 = AITHER_BC_SLIP_WALL;
    if (bc3 == AITHER_BC_VISCOUS_WALL) bc3 = AITHER_BC_SLIP_WALL;
  }
  for (int layer3 = 1; layer3 <= b.g; ++layer3) {
    for (int layer2 = 1; layer2 <= b.g; ++layer2) {
      const int pCellD2 = upper2 ? max2 + layer2 - 2 : 1 - layer2;
      const int gCellD2 = upper2 ? pCellD2 + 1 : pCellD2 - 1;
      const int pCellD3 = upper3 ? max3 + layer3 - 2 : 1 - layer3;
      const int gCellD3 = upper3 ? pCellD3 + 1 : pCellD3 - 1;
      int cg[3], cp2[3], cp3[3];
      DirIjk(dd, d1, gCellD2, gCellD3, cg);
      DirIjk(dd, d1, pCellD2, gCellD3, cp2);
      DirIjk(dd, d1, gCellD2, pCellD3, cp3);
      const long long ig = CellIdx(b, cg[0], cg[1], cg[2]);
      double from2[E::neq], from3[E::neq], ghost[E::neq];
      // plain loads: these cells may have been written earlier in this very loop
      for (int e = 0; e < E::neq; ++e) {
        from2[e] = b.state[e * b.fs + CellIdx(b, cp2[0], cp2[1], cp2[2])];
        from3[e] = b.state[e * b.fs + CellIdx(b, cp3[0], cp3[1], cp3[2])];
      }
      const bool wall2 = bc2 == AITHER_BC_SLIP_WALL && bc3 != AITHER_BC_SLIP_WALL;
      const bool wall3 = bc2 != AITHER_BC_SLIP_WALL && bc3 == AITHER_BC_SLIP_WALL;
      if (wall2 || wall3) {
        // the wall is extended into the edge cell: slip-wall reflection about the corner face
        int cf[3];
        if (wall2) DirIjk(dd, d1, cFaceD2_2, gCellD3, cf);
        else DirIjk(dd, d1, gCellD2, cFaceD3_3, cf);
        const int fd = wall2 ? fd2 : fd3;
        const long long fidx = CellIdx(b, cf[0], cf[1], cf[2]);
        double area[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) area[q] = b.fA[fd][q * b.fs + fidx];
        const int sx = wall2 ? s2 : s3;
        GhostState<NS, NT>(p.gas, wall2 ? from2 : from3, AITHER_BC_SLIP_WALL, area,
                           wall2 ? surf2 : surf3, bcs[surfs[sx].bcIndex], wall2 ? layer2 : layer3,
                           ghost);
      } else if (!VISCOUS || (bc2 == AITHER_BC_VISCOUS_WALL && bc3 == AITHER_BC_VISCOUS_WALL)) {
        if (layer2 == layer3) {
#pragma unroll
          for (int e = 0; e < E::neq; ++e) ghost[e] = 0.5 * (from2[e] + from3[e]);
        } else if (layer2 > layer3) {
#pragma unroll
          for (int e = 0; e < E::neq; ++e) ghost[e] = from3[e];
        } else {
#pragma unroll
          for (int e = 0; e < E::neq; ++e) ghost[e] = from2[e];
        }
      } else {
        continue;
      }
      for (int e = 0; e < E::neq; ++e) b.state[e * b.fs + ig] = ghost[e];
    }
  }
}

// ---- temperature and viscosity of every cell the stencils read (all but the corner ghosts) -----
// ref: src/procBlock.cpp:6171-6190 (UpdateAuxillaryVariables)
template <int NS, int NT>
__global__ void __launch_bounds__(256) AuxKernel(BlockDev b, Params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - b.g;
  const int j = blockIdx.y * blockDim.y + threadIdx.y - b.g;
  const int k = static_cast<int>(blockIdx.z) - b.g;
  if (i >= b.ni + b.g || j >= b.nj + b.g) return;
  const int out = (i < 0 || i >= b.ni) + (j < 0 || j >= b.nj) + (k < 0 || k >= b.nk);
  if (out == 3) return;
  const long long idx = CellIdx(b, i, j, k);
  double s[NS + 4 + NT];
  LoadCell<NS + 4 + NT>(b.state, b.fs, idx, s);
  const double t = Temperature<NS>(p.gas, s);
  b.temperature[idx] = t;
  b.viscosity[idx] = MixtureViscosity<NS>(p.tr, t, s);
}

// projected centre-to-centre distance across every face (geometry only; built once):
// ref: src/procBlock.cpp:6316-6341 (ProjC2CDist)
static __global__ void DistKernel(BlockDev b) {
  const int NI = b.ni + 2 * b.g, NJ = b.nj + 2 * b.g, NK = b.nk + 2 * b.g;
  const long long n = static_cast<long long>(NI) * NJ * NK;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c[3] = {static_cast<int>(t % NI) - b.g, static_cast<int>((t / NI) % NJ) - b.g,
                      static_cast<int>(t / (static_cast<long long>(NI) * NJ)) - b.g};
    const long long idx = CellIdx(b, c[0], c[1], c[2]);
    for (int d = 0; d < 3; ++d) {
      double dist = 1.0;
      if (c[d] > -b.g) {
        const long long lo = idx - Stride(b, d);
        double acc = 0.0;
        for (int q = 0; q < 3; ++q)
          acc += (b.center[q * b.fs + idx] - b.center[q * b.fs + lo]) * b.fA[d][q * b.fs + idx];
        dist = acc;
      }
      b.dist[d][idx] = dist;
    }
  }
}

// ---- K2/K3: face gradients + viscous flux ------------------------------------------------------
__device__ __forceinline__ void AreaVec(const BlockDev &b, int d, long long idx, double *v) {
  const double m = __ldg(b.fA[d] + 3 * b.fs + idx);
  v[0] = __ldg(b.fA[d] + idx) * m;  // unitVec3dMag::Vector(): unit * mag (vector3d.hpp:178)
  v[1] = __ldg(b.fA[d] + b.fs + idx) * m;
  v[2] = __ldg(b.fA[d] + 2 * b.fs + idx) * m;
}

// viscous flux times face area through face (idx = cell above the face) of direction D:
// out = {tau_x, tau_y, tau_z, tau.v + k grad T.n} |A|
template <int NS, int NT, int D>
__device__ __forceinline__ void ViscFaceFlux(const BlockDev &b, const Params &p, long long idx,
                                             double *out) {
  using E = Eq<NS, NT>;
  const long long sd = Stride(b, D);
  const long long st[3] = {1LL, static_cast<long long>(b.sj), b.sk};
  // areas of the control volume centred on the face (ref: src/procBlock.cpp:5190-5206)
  double al[3][3], au[3][3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double a0[3], a1[3];
    if (q == D) {
      AreaVec(b, D, idx, a0);
      AreaVec(b, D, idx + sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      AreaVec(b, D, idx - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    } else {
      AreaVec(b, q, idx + st[q], a0);
      AreaVec(b, q, idx + st[q] - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      AreaVec(b, q, idx, a0);
      AreaVec(b, q, idx - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    }
  }
  const double vol = 0.5 * (__ldg(b.vol + idx - sd) + __ldg(b.vol + idx));
  const double invVol = 1.0 / vol;
  // velocity (3) and temperature on the six faces of the control volume, then Green-Gauss
  // (ref: src/utility.cpp:59-175): grad(r, c) = sum_faces value_c * area_r / vol, i-, j-, k-pairs
  double grad[4][3];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double *f = c < 3 ? b.state + (NS + c) * b.fs : b.temperature;
    const double lo = __ldg(f + idx - sd), hi = __ldg(f + idx);
    double vl[3], vu[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (q == D) {
        vl[q] = lo;
        vu[q] = hi;
      } else {
        vu[q] = 0.25 * (lo + hi + __ldg(f + idx + st[q]) + __ldg(f + idx + st[q] - sd));
        vl[q] = 0.25 * (lo + hi + __ldg(f + idx - st[q]) + __ldg(f + idx - st[q] - sd));
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double t = vu[0] * au[0][r] - vl[0] * al[0][r] + vu[1] * au[1][r] - vl[1] * al[1][r] +
                       vu[2] * au[2][r] - vl[2] * al[2][r];
      grad[c][r] = t * invVol;
    }
  }
  // face state and viscosity: central or 4th-order central (ref: src/procBlock.cpp:1305-1346,
  // include/reconstruction.hpp:315-379)
  double fs_[E::neq], mu;
  if (p.viscRecon == 0) {
    const double w[2] = {__ldg(b.cw[D] + idx - sd), __ldg(b.cw[D] + idx)};
    double c[2];
    LagrangeCoeff<1>(w, 0, 0, c);
#pragma unroll
    for (int e = 0; e < E::neq; ++e)
      fs_[e] = c[0] * __ldg(b.state + e * b.fs + idx) + c[1] * __ldg(b.state + e * b.fs + idx - sd);
    mu = c[0] * __ldg(b.viscosity + idx) + c[1] * __ldg(b.viscosity + idx - sd);
  } else {
    const double w[4] = {__ldg(b.cw[D] + idx - 2 * sd), __ldg(b.cw[D] + idx - sd),
                         __ldg(b.cw[D] + idx), __ldg(b.cw[D] + idx + sd)};
    double c[4];
    LagrangeCoeff<3>(w, 1, 1, c);
#pragma unroll
    for (int e = 0; e < E::neq; ++e)
      fs_[e] = c[0] * __ldg(b.state + e * b.fs + idx - 2 * sd) +
               c[1] * __ldg(b.state + e * b.fs + idx - sd) + c[2] * __ldg(b.state + e * b.fs + idx) +
               c[3] * __ldg(b.state + e * b.fs + idx + sd);
    mu = c[0] * __ldg(b.viscosity + idx - 2 * sd) + c[1] * __ldg(b.viscosity + idx - sd) +
         c[2] * __ldg(b.viscosity + idx) + c[3] * __ldg(b.viscosity + idx + sd);
  }
  // viscousFlux::CalcFlux (src/viscousFlux.cpp:58-135), TauNormal (src/utility.cpp:425-437)
  double n[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) n[q] = __ldg(b.fA[D] + q * b.fs + idx);
  const double mag = __ldg(b.fA[D] + 3 * b.fs + idx);
  const double mus = p.tr.scaling * mu;
  const double lambda = 0.0 - (2.0 / 3.0) * (mus + 0.0);
  const double trace = grad[0][0] + grad[1][1] + grad[2][2];
  double tau[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // velGrad(r, c) = d u_c / d x_r = grad[c][r]; ((G + G^T) n)_r
    double mm = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) mm += (grad[c][r] + grad[r][c]) * n[c];
    tau[r] = lambda * trace * n[r] + (mus + 0.0) * mm;
  }
  const double t = Temperature<NS>(p.gas, fs_);
  const double kcond = EffectiveConductivity(p.tr, t);
  const double fe = (tau[0] * fs_[E::imx] + tau[1] * fs_[E::imy] + tau[2] * fs_[E::imz]) +
                    (kcond + 0.0) * (grad[3][0] * n[0] + grad[3][1] * n[1] + grad[3][2] * n[2]) + 0.0;
  out[0] = tau[0] * mag;
  out[1] = tau[1] * mag;
  out[2] = tau[2] * mag;
  out[3] = fe * mag;
}

// scratch: 4 doubles per face and direction. i-faces in vscr[0..3], j in [4..7], k in [8..11]
// (field stride fs, face indexed like its upper cell)
template <int NS, int NT>
__global__ void __launch_bounds__(256) ViscFaceKernel(BlockDev b, Params p, double *__restrict__ scrA,
                                                      double *__restrict__ scrB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i > b.ni || j > b.nj || k > b.nk) return;
  const long long idx = CellIdx(b, i, j, k);
  double f[4];
  if (j < b.nj && k < b.nk) {
    ViscFaceFlux<NS, NT, 0>(b, p, idx, f);
#pragma unroll
    for (int q = 0; q < 4; ++q) scrA[q * b.fs + idx] = f[q];
  }
  if (i < b.ni && k < b.nk) {
    ViscFaceFlux<NS, NT, 1>(b, p, idx, f);
    scrA[4 * b.fs + idx] = f[0];
#pragma unroll
    for (int q = 1; q < 4; ++q) scrB[(q - 1) * b.fs + idx] = f[q];
  }
  if (i < b.ni && j < b.nj) {
    ViscFaceFlux<NS, NT, 2>(b, p, idx, f);
    scrB[3 * b.fs + idx] = f[0];
    scrB[4 * b.fs + idx] = f[1];
    // the last two ride in the (otherwise idle) second half of the MUSCL-coefficient-free
    // matrix-residual field
    b.mres[0 * b.fs + idx] = f[2];
    b.mres[1 * b.fs + idx] = f[3];
  }
}

// per cell: residual += viscous fluxes (+lower, -upper; i, j, k), spectral radius and diagonal
// (ref: src/procBlock.cpp:1392-1493 and the J/K twins)
template <int NS, int NT>
__global__ void __launch_bounds__(256)
    ViscAccumKernel(BlockDev b, Params p, const double *__restrict__ scrA,
                    const double *__restrict__ scrB, int implicitScalar) {
  using E = Eq<NS, NT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  const long long st[3] = {1LL, static_cast<long long>(b.sj), b.sk};
  double s[E::neq];
  LoadCell<E::neq>(b.state, b.fs, idx, s);
  const double rho = SpeciesSum<NS>(s);
  const double gam = Gamma<NS>(p.gas, s);
  const double fac = ViscSpecFactor(p.tr, rho, gam, __ldg(b.viscosity + idx));
  const double vol = __ldg(b.vol + idx);
  double r[4], sr = b.specRad[idx], dg = implicitScalar ? b.diag[idx] : 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) r[q] = b.resid[(NS + q) * b.fs + idx];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double lo[4], hi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int slot = 4 * d + q;  // 0..4 scrA, 5..9 scrB, 10..11 mres
      const double *f = slot < 5 ? scrA + slot * b.fs
                                 : (slot < 10 ? scrB + (slot - 5) * b.fs : b.mres + (slot - 10) * b.fs);
      lo[q] = f[idx];
      hi[q] = f[idx + st[d]];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r[q] += lo[q];
      r[q] -= hi[q];
    }
    const double fMag = 0.5 * (__ldg(b.fA[d] + 3 * b.fs + idx) + __ldg(b.fA[d] + 3 * b.fs + idx + st[d]));
    const double vsr = fac * fMag * fMag / vol;
    sr += vsr * p.viscCFLCoeff;
    dg += 2.0 * vsr;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) b.resid[(NS + q) * b.fs + idx] = r[q];
  b.specRad[idx] = sr;
  if (implicitScalar) b.diag[idx] = dg;
}

// =============================================================================================
// RANS (k-omega Wilcox 2006 / SST 2003): viscous + turbulent face fluxes, cell averages of the
// face gradients / eddy viscosity / blending functions, turbulent spectral radii and the
// turbulence source terms, in ONE pass: a thread owns a cell and evaluates its six faces (each
// interior face is therefore evaluated twice, from identical inputs with identical code, so both
// owners see the same bits: owner-writes, no atomics, no scratch fields). Accumulation order per
// cell as the reference: i-lo, i-hi, j-lo, j-hi, k-lo, k-hi, then the source terms
// (ref: src/procBlock.cpp:1233-1497 and the J/K twins; CalcSrcTerms :5956-6025).
template <int NEQ>
struct FaceOut {
  double flux[NEQ];  // viscous flux * |A| (species rows are 0 for one species)
  double mut, f1, f2;
  double vg[9], kg[3], wg[3];
  double st[NEQ], mu;  // face state and laminar viscosity (thin-shear-layer Jacobian)
};

template <int NS, int NT, int D>
__device__ __forceinline__ void RansFace(const BlockDev &b, const Params &p, long long idx,
                                         FaceOut<NS + 4 + NT> &o, bool lowReWall = false,
                                         const double *wlv = nullptr, bool wallUpper = false) {
  using E = Eq<NS, NT>;
  constexpr int iw = E::it + (NT > 1 ? 1 : 0);
  constexpr int NG = NT > 0 ? 6 : 4;
  const long long sd = Stride(b, D);
  const long long st[3] = {1LL, static_cast<long long>(b.sj), b.sk};
  double al[3][3], au[3][3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double a0[3], a1[3];
    if (q == D) {
      AreaVec(b, D, idx, a0);
      AreaVec(b, D, idx + sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      AreaVec(b, D, idx - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    } else {
      AreaVec(b, q, idx + st[q], a0);
      AreaVec(b, q, idx + st[q] - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      AreaVec(b, q, idx, a0);
      AreaVec(b, q, idx - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    }
  }
  const double vol = 0.5 * (__ldg(b.vol + idx - sd) + __ldg(b.vol + idx));
  const double invVol = 1.0 / vol;
  // u, v, w, T (, k, omega) on the six faces of the control volume, then Green-Gauss
  // (ref: src/procBlock.cpp:5190-5352, src/utility.cpp:59-175)
  double grad[6][3];
#pragma unroll
  for (int r = 0; r < 3; ++r) grad[4][r] = grad[5][r] = 0.0;
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    const double *f = c < 3 ? b.state + (NS + c) * b.fs
                            : (c == 3 ? b.temperature : b.state + (E::it + c - 4) * b.fs);
    const double lo = __ldg(f + idx - sd), hi = __ldg(f + idx);
    double vl[3], vu[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (q == D) {
        vl[q] = lo;
        vu[q] = hi;
      } else {
        vu[q] = 0.25 * (lo + hi + __ldg(f + idx + st[q]) + __ldg(f + idx + st[q] - sd));
        vl[q] = 0.25 * (lo + hi + __ldg(f + idx - st[q]) + __ldg(f + idx - st[q] - sd));
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double t = vu[0] * au[0][r] - vl[0] * al[0][r] + vu[1] * au[1][r] - vl[1] * al[1][r] +
                       vu[2] * au[2][r] - vl[2] * al[2][r];
      grad[c][r] = t * invVol;
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) o.vg[3 * r + c] = grad[c][r];
    o.kg[r] = grad[4][r];
    o.wg[r] = grad[5][r];
  }
  // face state, viscosity and wall distance (ref: src/procBlock.cpp:1305-1351; turbulence
  // variables and the wall distance stay second order, include/reconstruction.hpp:359-379)
  double fs_[E::neq], mu, wDist = 0.0;
  {
    const double w2[2] = {__ldg(b.cw[D] + idx - sd), __ldg(b.cw[D] + idx)};
    double c2[2];
    LagrangeCoeff<1>(w2, 0, 0, c2);
    if (p.viscRecon == 0) {
#pragma unroll
      for (int e = 0; e < E::neq; ++e)
        fs_[e] = c2[0] * __ldg(b.state + e * b.fs + idx) + c2[1] * __ldg(b.state + e * b.fs + idx - sd);
      mu = c2[0] * __ldg(b.viscosity + idx) + c2[1] * __ldg(b.viscosity + idx - sd);
    } else {
      const double w[4] = {__ldg(b.cw[D] + idx - 2 * sd), w2[0], w2[1], __ldg(b.cw[D] + idx + sd)};
      double c[4];
      LagrangeCoeff<3>(w, 1, 1, c);
#pragma unroll
      for (int e = 0; e < NS + 4; ++e)
        fs_[e] = c[0] * __ldg(b.state + e * b.fs + idx - 2 * sd) +
                 c[1] * __ldg(b.state + e * b.fs + idx - sd) + c[2] * __ldg(b.state + e * b.fs + idx) +
                 c[3] * __ldg(b.state + e * b.fs + idx + sd);
#pragma unroll
      for (int e = NS + 4; e < E::neq; ++e)
        fs_[e] = c2[0] * __ldg(b.state + e * b.fs + idx) + c2[1] * __ldg(b.state + e * b.fs + idx - sd);
      mu = c[0] * __ldg(b.viscosity + idx - 2 * sd) + c[1] * __ldg(b.viscosity + idx - sd) +
           c[2] * __ldg(b.viscosity + idx) + c[3] * __ldg(b.viscosity + idx + sd);
    }
    if (b.wallDist) wDist = c2[0] * __ldg(b.wallDist + idx) + c2[1] * __ldg(b.wallDist + idx - sd);
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) fs_[E::it + t] = fmax(fs_[E::it + t], kTurbMin);  // LimitTurb
  if (wDist < 0.0 && wDist > -1.0e-10) wDist = 0.0;  // WALL_DIST_NEG_TOL
  const double rho = SpeciesSum<NS>(fs_);
  o.mut = o.f1 = o.f2 = 0.0;
  if (NT > 0)
    EddyViscAndBlending(p.tr.turbModel, p.tr.scaling, rho, fs_[E::it], fs_[iw], o.vg, o.kg, o.wg,
                        mu, wDist, &o.mut, &o.f1, &o.f2);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) o.st[e] = fs_[e];
  o.mu = mu;
  // viscousFlux::CalcFlux (src/viscousFlux.cpp:58-135), TauNormal (src/utility.cpp:425-437)
  double n[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) n[q] = __ldg(b.fA[D] + q * b.fs + idx);
  const double mag = __ldg(b.fA[D] + 3 * b.fs + idx);
  const double mus = p.tr.scaling * mu, muts = p.tr.scaling * o.mut;
  const double lambda = 0.0 - (2.0 / 3.0) * (mus + muts);
  const double trace = grad[0][0] + grad[1][1] + grad[2][2];
  double tau[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double mm = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) mm += (grad[c][r] + grad[r][c]) * n[c];
    tau[r] = lambda * trace * n[r] + (mus + muts) * mm;
  }
  const double t = Temperature<NS>(p.gas, fs_);
  const double kcond = MixtureEffConductivity<NS>(p.tr, t, fs_);
  // sutherland::TurbConductivity: mu_t cp / Pr_t (include/transport.hpp:132-137)
  const double kt = muts * Mixture<NS>(p.gas, fs_).cp / TurbPrandtl(p.tr.turbModel);
#pragma unroll
  for (int q = 0; q < NS; ++q) o.flux[q] = 0.0;
  double speciesEnthalpyTerm = 0.0;
  if (NS > 1 && !lowReWall) {
    // species diffusion D grad(Y_s).n with D = mu / Sc + mu_t / Sc_t (Sc_t = 0.7), rescaled so
    // that the positive and negative fluxes cancel, and the enthalpy it carries; a low-Re wall
    // face has none (ref: src/viscousFlux.cpp:84-105 vs :137-196; procBlock.cpp:5347-5376)
    const double dc = p.tr.schmidt > 0.0 ? mus / p.tr.schmidt + muts / 0.7 : 0.0;
    double fl[NS], posDiff = 0.0, negDiff = 0.0;
#pragma unroll
    for (int ss = 0; ss < NS; ++ss) {
      auto Y = [&](long long c) {
        double r = 0.0;
#pragma unroll
        for (int q = 0; q < NS; ++q) r += __ldg(b.state + q * b.fs + c);
        return __ldg(b.state + ss * b.fs + c) / r;
      };
      const double lo = Y(idx - sd), hi = Y(idx);
      double g3[3], vl[3], vu[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (q == D) {
          vl[q] = lo;
          vu[q] = hi;
        } else {
          vu[q] = 0.25 * (lo + hi + Y(idx + st[q]) + Y(idx + st[q] - sd));
          vl[q] = 0.25 * (lo + hi + Y(idx - st[q]) + Y(idx - st[q] - sd));
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double tt = vu[0] * au[0][r] - vl[0] * al[0][r] + vu[1] * au[1][r] - vl[1] * al[1][r] +
                          vu[2] * au[2][r] - vl[2] * al[2][r];
        g3[r] = tt * invVol;
      }
      fl[ss] = dc * Dot3(g3, n);
      negDiff -= fmin(fl[ss], 0.0);
      posDiff += fmax(fl[ss], 0.0);
    }
    const double posFac = posDiff > negDiff ? negDiff / posDiff : 1.0;
    const double negFac = negDiff > posDiff ? posDiff / negDiff : 1.0;
    const double vmag = sqrt(VelMagSq<NS>(fs_));
#pragma unroll
    for (int ss = 0; ss < NS; ++ss) {
      fl[ss] *= fl[ss] > 0.0 ? posFac : negFac;
      const double hs = (p.gas.hf[ss] + (p.gas.R[ss] * (p.gas.n[ss] + 1.0)) * t) + 0.5 * vmag * vmag;
      speciesEnthalpyTerm += fl[ss] * hs;
      o.flux[ss] = fl[ss] * mag;
    }
  }
  const double fe = (tau[0] * fs_[E::imx] + tau[1] * fs_[E::imy] + tau[2] * fs_[E::imz]) +
                    (kcond + kt) * (grad[3][0] * n[0] + grad[3][1] * n[1] + grad[3][2] * n[2]) +
                    speciesEnthalpyTerm;
  // k and omega diffusion; k-omega 2006 uses the unlimited eddy viscosity here
  // (ref: src/viscousFlux.cpp:117-134, include/turbulence.hpp:439)
  o.flux[E::imx] = tau[0] * mag;
  o.flux[E::imy] = tau[1] * mag;
  o.flux[E::imz] = tau[2] * mag;
  o.flux[E::ie] = fe * mag;
  if (NT > 0) {
    const double mutt = IsSst(p.tr.turbModel) ? muts : p.tr.scaling * (rho * fs_[E::it] / fs_[iw]);
    const double fk = (mus + TurbSigmaK(p.tr.turbModel, o.f1) * mutt) * Dot3(o.kg, n);
    const double fw = (mus + TurbSigmaW(p.tr.turbModel, o.f1) * mutt) * Dot3(o.wg, n);
    o.flux[E::it] = fk * mag;
    o.flux[iw] = fw * mag;
  }
  if (wlv != nullptr) {
    // boundary face on a wall-law wall: the gradients stand, everything else is prescribed by the
    // law -- wall state (src/wallData.cpp:299-313), wall viscosities, f1 = f2 = 1
    // (src/procBlock.cpp:1286-1300) and viscousFlux::CalcWallLawFlux (src/viscousFlux.cpp:213-249)
    const double invScaling = 1.0 / p.tr.scaling;
    o.f1 = 1.0;
    o.f2 = 1.0;
    o.mu = wlv[kWvMu] * invScaling;
    o.mut = wlv[kWvMut] * invScaling;
    const long long ca = wallUpper ? idx - sd : idx;  // wall mass fractions: the adjacent cell's
    double rhoA = 0.0, pW = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) rhoA += __ldg(b.state + q * b.fs + ca);
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      o.st[q] = (__ldg(b.state + q * b.fs + ca) / rhoA) * wlv[kWvRho];
      pW += o.st[q] * p.gas.R[q];
      o.flux[q] = 0.0;
    }
    const double vw[3] = {wlv[kWvVelWall], wlv[kWvVelWall + 1], wlv[kWvVelWall + 2]};
    o.st[E::imx] = vw[0];
    o.st[E::imy] = vw[1];
    o.st[E::imz] = vw[2];
    o.st[E::ie] = pW * wlv[kWvT];
    const double tw[3] = {wlv[kWvTau], wlv[kWvTau + 1], wlv[kWvTau + 2]};
    o.flux[E::imx] = tw[0] * mag;
    o.flux[E::imy] = tw[1] * mag;
    o.flux[E::imz] = tw[2] * mag;
    o.flux[E::ie] = ((tw[0] * vw[0] + tw[1] * vw[1] + tw[2] * vw[2]) + wlv[kWvHeatFlux]) * mag;
    if (NT > 0) {
      o.st[E::it] = wlv[kWvTke];
      o.st[iw] = wlv[kWvSdr];
      const double fk = (wlv[kWvMu] + TurbSigmaK(p.tr.turbModel, 1.0) * wlv[kWvMut]) * Dot3(o.kg, n);
      const double fw = (wlv[kWvMu] + TurbSigmaW(p.tr.turbModel, 1.0) * wlv[kWvMut]) * Dot3(o.wg, n);
      o.flux[E::it] = fk * mag;
      o.flux[iw] = fw * mag;
    }
  }
}

// ---- cell averages of the face pressure (and velocity) gradients for the non-reflecting BCs ----
// The reference's viscous flux loops leave pressureGrad_ / velocityGrad_ behind as 1/6 of the six
// face gradients of every cell (src/procBlock.cpp:1396-1452), and the non-reflecting inlet /
// outlet of the NEXT iteration reads them in the boundary-adjacent cells (:2513-2514). Runs
// without such a BC never launch this kernel. Same Green-Gauss control volume and summation
// order as RansFace; `withVel`: also the velocity gradient (the laminar scalar path does not
// produce it otherwise).
template <int D>
__device__ __forceinline__ void FaceControlVolume(const BlockDev &b, long long idx, double al[3][3],
                                                  double au[3][3], double *invVol) {
  const long long sd = Stride(b, D);
  const long long st[3] = {1LL, static_cast<long long>(b.sj), b.sk};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double a0[3], a1[3];
    if (q == D) {
      AreaVec(b, D, idx, a0);
      AreaVec(b, D, idx + sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      AreaVec(b, D, idx - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    } else {
      AreaVec(b, q, idx + st[q], a0);
      AreaVec(b, q, idx + st[q] - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      AreaVec(b, q, idx, a0);
      AreaVec(b, q, idx - sd, a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    }
  }
  *invVol = 1.0 / (0.5 * (__ldg(b.vol + idx - sd) + __ldg(b.vol + idx)));
}
template <int D>
__device__ __forceinline__ void FaceGreenGauss(const BlockDev &b, const double *f, long long idx,
                                               const double al[3][3], const double au[3][3],
                                               double invVol, double *g3) {
  const long long sd = Stride(b, D);
  const long long st[3] = {1LL, static_cast<long long>(b.sj), b.sk};
  const double lo = f[idx - sd], hi = f[idx];
  double vl[3], vu[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (q == D) {
      vl[q] = lo;
      vu[q] = hi;
    } else {
      vu[q] = 0.25 * (lo + hi + f[idx + st[q]] + f[idx + st[q] - sd]);
      vl[q] = 0.25 * (lo + hi + f[idx - st[q]] + f[idx - st[q] - sd]);
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double t = vu[0] * au[0][r] - vl[0] * al[0][r] + vu[1] * au[1][r] - vl[1] * al[1][r] +
                     vu[2] * au[2][r] - vl[2] * al[2][r];
    g3[r] = t * invVol;
  }
}
template <int NS, int D>
__device__ __forceinline__ void CellGradDir(const BlockDev &b, long long idx, bool withVel,
                                            double *pg, double *vg) {
  constexpr double sixth = 1.0 / 6.0;
  const long long sd = Stride(b, D);
#pragma unroll
  for (int side = 0; side < 2; ++side) {  // lower face, then upper face (reference order)
    const long long fidx = idx + side * sd;
    double al[3][3], au[3][3], invVol, g3[3];
    FaceControlVolume<D>(b, fidx, al, au, &invVol);
    FaceGreenGauss<D>(b, b.state + (NS + 3) * b.fs, fidx, al, au, invVol, g3);
#pragma unroll
    for (int r = 0; r < 3; ++r) pg[r] += sixth * g3[r];
    if (withVel) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        FaceGreenGauss<D>(b, b.state + (NS + c) * b.fs, fidx, al, au, invVol, g3);
#pragma unroll
        for (int r = 0; r < 3; ++r) vg[3 * r + c] += sixth * g3[r];
      }
    }
  }
}
template <int NS, int NT>
__global__ void __launch_bounds__(128) CellGradKernel(BlockDev b, int withVel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double pg[3] = {0.0, 0.0, 0.0}, vg[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) vg[q] = 0.0;
  CellGradDir<NS, 0>(b, idx, withVel != 0, pg, vg);
  CellGradDir<NS, 1>(b, idx, withVel != 0, pg, vg);
  CellGradDir<NS, 2>(b, idx, withVel != 0, pg, vg);
#pragma unroll
  for (int q = 0; q < 3; ++q) b.pressGrad[q * b.fs + idx] = pg[q];
  if (withVel) {
#pragma unroll
    for (int q = 0; q < 9; ++q) b.velGrad[q * b.fs + idx] = vg[q];
  }
}

template <int NS, int NT>
struct RansAcc {
  double r[NS + 4 + NT];
  double sr, srT, dg, dgT;
  double mut, f1, f2, vg[9], kg[3], wg[3];
};

template <int NS, int NT, int D, bool BLOCK>
__device__ __forceinline__ void RansAccumulateDir(const BlockDev &b, const Params &p, long long idx,
                                                  const double *s, double visc, double vol,
                                                  RansAcc<NS, NT> &a, double *dblk, bool wallLo,
                                                  bool wallHi, const double *wlvLo = nullptr,
                                                  const double *wlvHi = nullptr) {
  using E = Eq<NS, NT>;
  constexpr double sixth = 1.0 / 6.0;
  constexpr int iw = E::it + (NT > 1 ? 1 : 0);
  const long long sd = Stride(b, D);
  FaceOut<E::neq> f;
  // lower face: this cell is the face's upper cell (ref: :1432-1493)
  RansFace<NS, NT, D>(b, p, idx, f, wallLo, wlvLo, false);
#pragma unroll
  for (int e = (NS > 1 ? 0 : NS); e < E::neq; ++e) a.r[e] += f.flux[e];
  const double mutLo = f.mut, f1Lo = f.f1;
#pragma unroll
  for (int q = 0; q < 9; ++q) a.vg[q] += sixth * f.vg[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    a.kg[q] += sixth * f.kg[q];
    a.wg[q] += sixth * f.wg[q];
  }
  a.mut += sixth * f.mut;
  a.f1 += sixth * f.f1;
  a.f2 += sixth * f.f2;
  {
    const double fMag = 0.5 * (__ldg(b.fA[D] + 3 * b.fs + idx) + __ldg(b.fA[D] + 3 * b.fs + idx + sd));
    const double rho = SpeciesSum<NS>(s);
    const double length = fMag * fMag / vol;
    const double vsr = ViscSpecFactor(p.tr, rho, Gamma<NS>(p.gas, s), visc, mutLo) * length;
    const double tvsr = NT > 0 ? TurbViscSpecFactor(p.tr.turbModel, p.tr.scaling, rho, s[E::it],
                                                    s[iw], visc, mutLo, f1Lo) * length
                               : 0.0;
    a.sr += vsr * p.viscCFLCoeff;
    a.srT += tvsr * p.viscCFLCoeff;
    a.dg += 2.0 * vsr;
    a.dgT += 2.0 * tvsr;
  }
  if constexpr (BLOCK) {  // + dFv/dU of the lower face (left = false); ref: src/procBlock.cpp:1481-1489
    double fa[4], J[Blk<NS, NT>::n];
#pragma unroll
    for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[D] + q * b.fs + idx);
    ApproxTslJacobian<NS, NT>(p.gas, p.tr, f.st, f.mu, f.mut, f.f1, fa, __ldg(b.dist[D] + idx), false,
                              f.vg, J);
#pragma unroll
    for (int q = 0; q < Blk<NS, NT>::n; ++q) dblk[q] += J[q];
  }
  // upper face: this cell is the face's lower cell (ref: :1392-1429)
  RansFace<NS, NT, D>(b, p, idx + sd, f, wallHi, wlvHi, true);
#pragma unroll
  for (int e = (NS > 1 ? 0 : NS); e < E::neq; ++e) a.r[e] -= f.flux[e];
  if constexpr (BLOCK) {  // - dFv/dU of the upper face (left = true); ref: src/procBlock.cpp:1420-1428
    double fa[4], J[Blk<NS, NT>::n];
#pragma unroll
    for (int q = 0; q < 4; ++q) fa[q] = __ldg(b.fA[D] + q * b.fs + idx + sd);
    ApproxTslJacobian<NS, NT>(p.gas, p.tr, f.st, f.mu, f.mut, f.f1, fa, __ldg(b.dist[D] + idx + sd),
                              true, f.vg, J);
#pragma unroll
    for (int q = 0; q < Blk<NS, NT>::n; ++q) dblk[q] -= J[q];
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) a.vg[q] += sixth * f.vg[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    a.kg[q] += sixth * f.kg[q];
    a.wg[q] += sixth * f.wg[q];
  }
  a.mut += sixth * f.mut;
  a.f1 += sixth * f.f1;
  a.f2 += sixth * f.f2;
}

template <int NS, int NT, bool BLOCK>
__global__ void __launch_bounds__(128)
    RansCellKernel(BlockDev b, Params p, int implicitScalar, const EdgeSurf *__restrict__ surfs,
                   int nsurf) {
  using E = Eq<NS, NT>;
  using B = Blk<NS, NT>;
  constexpr int iw = E::it + (NT > 1 ? 1 : 0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= b.ni || j >= b.nj) return;
  const long long idx = CellIdx(b, i, j, k);
  double s[E::neq];
  LoadCell<E::neq>(b.state, b.fs, idx, s);
  const double visc = __ldg(b.viscosity + idx);
  const double vol = __ldg(b.vol + idx);
  RansAcc<NS, NT> a;
  double dblk[BLOCK ? B::n : 1];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) a.r[e] = b.resid[e * b.fs + idx];
  a.sr = b.specRad[idx];
  a.srT = b.specRad[b.fs + idx];
  a.dg = implicitScalar ? b.diag[idx] : 0.0;
  a.dgT = (implicitScalar && NT > 0) ? b.diag[b.fs + idx] : 0.0;
  if (BLOCK) {
#pragma unroll
    for (int q = 0; q < B::n; ++q) dblk[q] = b.diag[q * b.fs + idx];
  }
  a.mut = a.f1 = a.f2 = 0.0;
#pragma unroll
  for (int q = 0; q < 9; ++q) a.vg[q] = 0.0;
#pragma unroll
  for (int q = 0; q < 3; ++q) a.kg[q] = a.wg[q] = 0.0;
  // multi-species: boundary faces on a viscous wall carry no species diffusion
  // and a wall-law wall whose y+ stayed >= 10 prescribes the whole flux of its faces
  bool wl[3] = {false, false, false}, wh[3] = {false, false, false};
  const double *vl[3] = {nullptr, nullptr, nullptr}, *vh[3] = {nullptr, nullptr, nullptr};
  if (NS > 1 || b.wallVars != nullptr) {
    const int c[3] = {i, j, k}, nd[3] = {b.ni, b.nj, b.nk};
    auto wallRecord = [&](int sf, int d) -> const double * {
      // record index as written by ViscousWallKernel: d1 + n1 * d2 with the reference's
      // direction cycling
      if (b.wallVars == nullptr || surfs[sf].faceBase < 0) return nullptr;
      const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      const int n1 = surfs[sf].hi[d1] - surfs[sf].lo[d1];
      const double *w = b.wallVars + kWallVarsStride * (static_cast<long long>(surfs[sf].faceBase) +
                                                        (c[d1] - surfs[sf].lo[d1]) +
                                                        static_cast<long long>(n1) * (c[d2] - surfs[sf].lo[d2]));
      return w[kWvYplus] < 10.0 ? nullptr : w;
    };
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (c[d] == 0) {
        const int sf = FindSurface(surfs, nsurf, c, 2 * d + 1);
        if (sf >= 0 && surfs[sf].type == AITHER_BC_VISCOUS_WALL) {
          vl[d] = wallRecord(sf, d);
          wl[d] = vl[d] == nullptr;
        }
      }
      if (c[d] == nd[d] - 1) {
        int cu[3] = {i, j, k};
        cu[d] += 1;
        const int sf = FindSurface(surfs, nsurf, cu, 2 * d + 2);
        if (sf >= 0 && surfs[sf].type == AITHER_BC_VISCOUS_WALL) {
          vh[d] = wallRecord(sf, d);
          wh[d] = vh[d] == nullptr;
        }
      }
    }
  }
  RansAccumulateDir<NS, NT, 0, BLOCK>(b, p, idx, s, visc, vol, a, dblk, wl[0], wh[0], vl[0], vh[0]);
  RansAccumulateDir<NS, NT, 1, BLOCK>(b, p, idx, s, visc, vol, a, dblk, wl[1], wh[1], vl[1], vh[1]);
  RansAccumulateDir<NS, NT, 2, BLOCK>(b, p, idx, s, visc, vol, a, dblk, wl[2], wh[2], vl[2], vh[2]);
  if (NT > 0) {
    // source terms (ref: src/procBlock.cpp:5956-6025, src/source.cpp:64-82)
    double src[2], beta = 0.0;
    const double rho = SpeciesSum<NS>(s);
    TurbSource(p.tr.turbModel, p.tr.scaling, rho, s[E::it], s[iw], a.vg, a.kg, a.wg, a.mut, a.f1,
               src, &beta);
    const double turbSpecRad = TurbSrcSpecRad(p.tr.scaling, s[iw], vol);
    a.srT -= turbSpecRad;
    a.dgT -= turbSpecRad;
    a.r[E::it] -= src[0] * vol;
    a.r[iw] -= src[1] * vol;
    if (BLOCK) {  // TurbSrcJac (src/turbulence.cpp:445-458, :706-720)
      const double invScaling = 1.0 / p.tr.scaling;
      dblk[B::nf] -= -2.0 * kw::betaStar * s[iw] * vol * invScaling;
      dblk[B::nf + NT * NT - 1] -= -2.0 * beta * s[iw] * vol * invScaling;
    }
  }
#pragma unroll
  for (int e = (NS > 1 ? 0 : NS); e < E::neq; ++e) b.resid[e * b.fs + idx] = a.r[e];
  b.specRad[idx] = a.sr;
  b.specRad[b.fs + idx] = a.srT;
  if (implicitScalar) {
    b.diag[idx] = a.dg;
    if (NT > 0) b.diag[b.fs + idx] = a.dgT;
  }
  if (BLOCK) {
#pragma unroll
    for (int q = 0; q < B::n; ++q) b.diag[q * b.fs + idx] = dblk[q];
  }
  if (NT > 0) {
    b.eddyVisc[idx] = a.mut;
    b.f1[idx] = a.f1;
    b.f2[idx] = a.f2;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      b.tkeGrad[q * b.fs + idx] = a.kg[q];
      b.omegaGrad[q * b.fs + idx] = a.wg[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) b.velGrad[q * b.fs + idx] = a.vg[q];
}

}  // namespace aither
