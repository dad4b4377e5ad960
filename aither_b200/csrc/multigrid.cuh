// multigrid.cuh -- transfer operators between two grid levels. Every level
// is a handle of its own (aither_gpu_create with the coarse blocks); the full-approximation-
// storage cycle is composed by the caller from the per-level phases and these operators, as the
// reference composes it from gridLevel methods (mgSolution::CycleAtLevel, src/mgSolution.cpp:160-207).
//
// Reference: gridLevel::Restriction / Prolongation (src/gridLevel.cpp:538-611), BlockRestriction
// (include/procBlock.hpp:637-690), BlockProlongation (include/gridLevel.hpp:159-214),
// ConvertCellToNode / TrilinearInterp (include/utility.hpp:186-372).
//
// Device design: gather form, owner-writes, no atomics. The reference scatters fine cells into
// their coarse cell in k, j, i order; here a coarse cell sums its (at most eight) children in that
// same order from a child list built once on the host, so the sums round identically. The
// forcing term f = (A x - b) + restricted fine matrix residual is folded into the coarse level's
// right-hand side field (b := b + f): every sweep kernel and the matrix residual read b only, and
// the reference evaluates (b + f) + off-diagonal in that association (src/linearSolver.cpp:503).
// Fields are equation-count agnostic here (runtime neq): these kernels are HBM-trivial (one pass
// over a level that is 8x smaller than the one above).
#pragma once
#include "layout.cuh"

namespace aither {

constexpr int kMaxChildren = 8;

// coarse = sum over children of w * fine (w = volume weight, or 1 when volFac == nullptr);
// srcPadded / dstPadded: the field is indexed with the ghost-padded cell index (state, update) or
// with it too (matrix residual): all device fields share the padded layout
static __global__ void RestrictKernel(BlockDev f, BlockDev c, const int *__restrict__ children,
                                      const double *__restrict__ volFac, const double *src,
                                      double *dst, int neq, bool accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= c.ni || j >= c.nj) return;
  const long long pc = i + static_cast<long long>(c.ni) * (j + static_cast<long long>(c.nj) * k);
  const long long cidx = CellIdx(c, i, j, k);
  for (int e = 0; e < neq; ++e) {
    double acc = 0.0;
    for (int q = 0; q < kMaxChildren; ++q) {
      const int pf = children[kMaxChildren * pc + q];
      if (pf < 0) break;
      const int fi = pf % f.ni, fj = (pf / f.ni) % f.nj, fk = pf / (f.ni * f.nj);
      const double v = src[e * f.fs + CellIdx(f, fi, fj, fk)];
      acc = acc + (volFac != nullptr ? volFac[pf] * v : v);
    }
    if (accumulate) dst[e * c.fs + cidx] = dst[e * c.fs + cidx] + acc;
    else dst[e * c.fs + cidx] = acc;
  }
}

// forcing folded into the right-hand side: rhs := rhs + ((A x - b) + sum of the children's matrix
// residual); mres of the coarse level holds 0 - (A x - b) at this point
static __global__ void ForcingKernel(BlockDev f, BlockDev c, const int *__restrict__ children,
                                     int neq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= c.ni || j >= c.nj) return;
  const long long pc = i + static_cast<long long>(c.ni) * (j + static_cast<long long>(c.nj) * k);
  const long long cidx = CellIdx(c, i, j, k);
  for (int e = 0; e < neq; ++e) {
    double r = 0.0;
    for (int q = 0; q < kMaxChildren; ++q) {
      const int pf = children[kMaxChildren * pc + q];
      if (pf < 0) break;
      const int fi = pf % f.ni, fj = (pf / f.ni) % f.nj, fk = pf / (f.ni * f.nj);
      r = r + f.mres[e * f.fs + CellIdx(f, fi, fj, fk)];
    }
    const double axmb = 0.0 - c.mres[e * c.fs + cidx];
    const double forcing = axmb + r;
    c.rhs[e * c.fs + cidx] = c.rhs[e * c.fs + cidx] + forcing;
  }
}

static __global__ void AxpyFieldKernel(double *x, const double *y, double a, long long n) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x)
    x[t] = x[t] + a * y[t];
}

// node values of the coarse update: sum of the physical cells around the node in k, j, i order,
// times 1 at the block corners, 1/2 on the block edges, 1/8 elsewhere
// (ConvertCellToNode(coarse, ignoreEdge = true, ignoreGhosts = true))
static __global__ void NodeKernel(BlockDev c, const double *__restrict__ x, double *node, int neq) {
  const int n0 = c.ni + 1, n1 = c.nj + 1, n2 = c.nk + 1;
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y * blockDim.y + threadIdx.y;
  const int K = blockIdx.z;
  if (I >= n0 || J >= n1 || K >= n2) return;
  const bool ei = I == 0 || I == n0 - 1, ej = J == 0 || J == n1 - 1, ek = K == 0 || K == n2 - 1;
  const double fac = (ei && ej && ek) ? 1.0 : ((static_cast<int>(ei) + ej + ek == 2) ? 1.0 / 2.0 : 1.0 / 8.0);
  const long long nidx = I + static_cast<long long>(n0) * (J + static_cast<long long>(n1) * K);
  const long long nn = static_cast<long long>(n0) * n1 * n2;
  for (int e = 0; e < neq; ++e) {
    double acc = 0.0;
    for (int kk = K - 1; kk <= K; ++kk)
      for (int jj = J - 1; jj <= J; ++jj)
        for (int ii = I - 1; ii <= I; ++ii) {
          if (ii < 0 || jj < 0 || kk < 0 || ii >= c.ni || jj >= c.nj || kk >= c.nk) continue;
          acc += x[e * c.fs + CellIdx(c, ii, jj, kk)];
        }
    node[e * nn + nidx] = acc * fac;
  }
}

// trilinear interpolation of the coarse node values at every fine cell centre, added to the fine
// update (BlockProlongation + linearSolver::AddToUpdate)
static __global__ void ProlongKernel(BlockDev f, BlockDev c, const int *__restrict__ toCoarse,
                                     const double *__restrict__ coef,
                                     const double *__restrict__ node, double *x, int neq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= f.ni || j >= f.nj) return;
  const long long pf = i + static_cast<long long>(f.ni) * (j + static_cast<long long>(f.nj) * k);
  const int ci = toCoarse[3 * pf], cj = toCoarse[3 * pf + 1], ck = toCoarse[3 * pf + 2];
  const double *co = coef + 7 * pf;
  const int n0 = c.ni + 1, n1 = c.nj + 1, n2 = c.nk + 1;
  const long long nn = static_cast<long long>(n0) * n1 * n2;
  auto N = [&](int a, int b2, int d) {
    return (ci + a) + static_cast<long long>(n0) * ((cj + b2) + static_cast<long long>(n1) * (ck + d));
  };
  const long long fidx = CellIdx(f, i, j, k);
  for (int e = 0; e < neq; ++e) {
    const double *nd = node + e * nn;
    const double d0 = nd[N(0, 0, 0)], d1 = nd[N(1, 0, 0)], d2 = nd[N(0, 1, 0)], d3 = nd[N(1, 1, 0)],
                 d4 = nd[N(0, 0, 1)], d5 = nd[N(1, 0, 1)], d6 = nd[N(0, 1, 1)], d7 = nd[N(1, 1, 1)];
    const double d04 = (1.0 - co[0]) * d0 + co[0] * d4;
    const double d15 = (1.0 - co[1]) * d1 + co[1] * d5;
    const double d26 = (1.0 - co[2]) * d2 + co[2] * d6;
    const double d37 = (1.0 - co[3]) * d3 + co[3] * d7;
    const double d0415 = (1.0 - co[4]) * d04 + co[4] * d15;
    const double d2637 = (1.0 - co[5]) * d26 + co[5] * d37;
    x[e * f.fs + fidx] += (1.0 - co[6]) * d0415 + co[6] * d2637;
  }
}

}  // namespace aither
