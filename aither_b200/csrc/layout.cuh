// layout.cuh -- device data layout of one block.
//
// HBM layout (differs from the reference on purpose): structure-of-arrays, one plane-strided
// scalar field per variable, i fastest, with
//   * a left pad LP (multiple of 16 doubles = 128 B) so that physical cell i = 0 of every row
//     starts on a 128-byte line and warp-wide loads of 32 consecutive cells are fully coalesced;
//   * an i-pitch rounded up to 16 doubles;
//   * ONE index function for every cell- and face-centred field: all fields are allocated with
//     (nj + 2g + 1) rows and (nk + 2g + 1) planes so that face arrays (one longer in their own
//     direction; ref: src/procBlock.cpp:117-124) share the indexing of cell arrays.
// Face convention as in the reference (src/procBlock.cpp:371-376): face (i,j,k) of direction d
// lies between cell (i,j,k)-e_d and cell (i,j,k); its area vector points to increasing index.
#pragma once
#include <stdint.h>

namespace aither {

struct BlockDev {
  int ni, nj, nk, g;
  int lp;            // left pad in doubles
  int sj;            // j stride (i pitch), doubles
  long long sk;      // k stride
  long long fs;      // field stride = allocated doubles per scalar field
  int parentBlock;
  // fields (each `fs` doubles per component)
  double *state;     // neq   primitive, ghosts valid
  double *stateAlt;  // neq   the state the fused matrix-residual pass advances into (inviscid
                     //       one-species scalar-diagonal runs, else null); swapped with `state`
  double *consN;     // neq   U^n
  double *consNm1;   // neq   U^(n-1) (bdf2 only, else null)
  double *resid;     // neq   residual R
  double *rhs;       // neq   b = -R/theta + time terms
  double *x;         // neq   update (current)
  double *xalt;      // neq   update (ping-pong partner for DPLUR)
  double *mres;      // neq   matrix residual (kept only when requested)
  double *specRad;   // 2     {flow, turb}
  double *dt;        // 1
  double *diag;      // asz   main diagonal D
  double *dinv;      // asz   D^-1
  double *vol;       // 1
  double *cw[3];     // 1     cell widths along i, j, k
  double *mc[3];     // 2     MUSCL grid ratios along i, j, k: {2w/(w+w_lower), 2w/(w+w_upper)}
  double *fA[3];     // 4     face areas {nx, ny, nz, |A|} for i-, j-, k-faces
  double *center;    // 3
  // viscous runs only (else null)
  double *temperature;  // 1   ghosts valid (all but corner cells)
  double *viscosity;    // 1   laminar viscosity, normalised by mu_ref; ghosts valid
  double *wallDist;     // 1   (null when the caller gave none)
  double *dist[3];      // 1   projected centre-to-centre distance across i-, j-, k-faces
  // RANS runs only (else null): cell averages of the six face values
  // (ref: src/procBlock.cpp:1396-1452); eddyVisc, f1, f2 are contiguous fields (one halo exchange)
  double *eddyVisc;     // 1   ghosts valid across connections only
  double *f1, *f2;      // 1
  double *velGrad;      // 9   velGrad[3 r + c] = d u_c / d x_r
  double *tkeGrad;      // 3
  double *omegaGrad;    // 3
  // runs with non-reflecting BCs only (else null): cell average of the six face pressure gradients
  double *pressGrad;    // 3
  // wall-law runs only (else null): kWallVarsStride doubles per boundary face of the block
  // (walllaw.cuh), face index = SurfDev::faceOffset / g + d1 + n1 * d2
  double *wallVars;
  // per boundary face: 1 if the neighbour across that block face contributes to the implicit
  // off-diagonals, i.e. the face belongs to a connection (interblock / periodic) boundary
  // (ref: src/procBlock.cpp:1064,1115; include/boundaryConditions.hpp:287-293).
  // index [surf-1][d1 + n1 * d2] with the reference's direction cycling (i: j,k; j: k,i; k: i,j)
  const uint8_t *connFace[6];
};

__host__ __device__ __forceinline__ long long CellIdx(const BlockDev &b, int i, int j, int k) {
  return static_cast<long long>(i + b.lp) + static_cast<long long>(j + b.g) * b.sj +
         static_cast<long long>(k + b.g) * b.sk;
}
__host__ __device__ __forceinline__ long long Stride(const BlockDev &b, int d) {
  return d == 0 ? 1LL : (d == 1 ? static_cast<long long>(b.sj) : b.sk);
}

}  // namespace aither
