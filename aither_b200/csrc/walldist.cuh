// walldist.cuh -- distance of every cell centre to the nearest viscous-wall face centre.
//
// ref: src/procBlock.cpp:6030-6107 (procBlock::CalcWallDistance), src/kdtree.cpp:123-225
// (kdtree::NearestNeighbor: the smallest vector3d::DistSq, include/vector3d.hpp:366-371, returned
// as its square root), src/main.cpp:144,191-201.
//
// The reference builds a k-d tree because a CPU core cannot afford cells x faces distance
// evaluations. Here the exhaustive search is the simple and the fast way: a tile of wall points
// sits in shared memory, every thread keeps the running minimum of its cell -- three subtractions
// and three multiply-adds per pair, no tree to build, no divergence. The minimum of the squared
// distances is the tree's minimum, so the result is the reference's.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"

namespace aither {

constexpr int kWallTile = 1024;

static __global__ void __launch_bounds__(256)
    WallDistKernel(BlockDev b, const double *__restrict__ pts, long long nPts) {
  __shared__ double sp[3 * kWallTile];
  const long long nCells = static_cast<long long>(b.ni) * b.nj * b.nk;
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const bool valid = t < nCells;
  const int i = valid ? static_cast<int>(t % b.ni) : 0;
  const int j = valid ? static_cast<int>((t / b.ni) % b.nj) : 0;
  const int k = valid ? static_cast<int>(t / (static_cast<long long>(b.ni) * b.nj)) : 0;
  const long long idx = CellIdx(b, i, j, k);
  const double cx = b.center[idx], cy = b.center[b.fs + idx], cz = b.center[2 * b.fs + idx];
  double best = 1.7976931348623157e308;  // std::numeric_limits<double>::max(), kdtree.cpp:217
  for (long long base = 0; base < nPts; base += kWallTile) {
    const int n = static_cast<int>(min(static_cast<long long>(kWallTile), nPts - base));
    __syncthreads();
    for (int q = threadIdx.x; q < 3 * n; q += blockDim.x) sp[q] = pts[3 * base + q];
    __syncthreads();
    for (int q = 0; q < n; ++q) {
      const double dx = cx - sp[3 * q], dy = cy - sp[3 * q + 1], dz = cz - sp[3 * q + 2];
      const double d2 = dx * dx + dy * dy + dz * dz;  // vector3d::MagSq
      best = d2 < best ? d2 : best;
    }
  }
  if (valid) b.wallDist[idx] = sqrt(best);
}

// ghost cells of one boundary surface (not the edge ghost cells): across a viscous wall minus the
// mirrored interior value, so that the wall distance at the wall face is zero; elsewhere the
// value of the first interior cell (ref src/procBlock.cpp:6045-6104)
static __global__ void WallDistGhostKernel(BlockDev b, int d3, int isLower, int isWall, int lo1, int n1,
                                    int lo2, int n2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2 * b.g) return;
  const int d1 = (d3 + 1) % 3, d2 = (d3 + 2) % 3;
  const int nd[3] = {b.ni, b.nj, b.nk};
  const int layer = t / (n1 * n2) + 1, r = t % (n1 * n2);
  int cg[3], ci[3];
  cg[d1] = ci[d1] = lo1 + r % n1;
  cg[d2] = ci[d2] = lo2 + r / n1;
  cg[d3] = isLower ? -layer : nd[d3] + layer - 1;
  ci[d3] = isLower ? (isWall ? layer - 1 : 0) : (isWall ? nd[d3] - layer : nd[d3] - 1);
  const double v = b.wallDist[CellIdx(b, ci[0], ci[1], ci[2])];
  b.wallDist[CellIdx(b, cg[0], cg[1], cg[2])] = isWall ? -1.0 * v : v;
}

}  // namespace aither
