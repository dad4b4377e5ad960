// probe: which 4-D float64 tensor-map shapes the TMA accepts (debugging aid, not product code)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../aither_b200/csrc/tma.cuh"
using namespace aither;
__global__ void Probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int n,
                      double *out) {
  extern __shared__ __align__(128) double sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 8192);
  if (threadIdx.x == 0) {
    MbarInit(bar, 1);
    MbarInitFence();
    MbarExpectTx(bar, n * 8);
    TmaLoad4D(sm, &map, c0, c1, c2, c3, bar);
  }
  __syncthreads();
  MbarWait(bar, 0);
  for (int q = threadIdx.x; q < n; q += blockDim.x) out[q] = sm[q];
}
int main(int argc, char **argv) {
  const int bx = atoi(argv[1]), by = atoi(argv[2]), nf = atoi(argv[3]), c0 = atoi(argv[4]);
  BlockDev b{};
  b.ni = 40; b.nj = 20; b.nk = 6; b.g = 2; b.lp = 16;
  b.sj = ((b.lp + b.ni + b.g + 1 + 15) / 16) * 16;
  b.sk = (long long)b.sj * (b.nj + 2 * b.g + 1);
  b.fs = ((b.sk * (b.nk + 2 * b.g + 1) + 15) / 16) * 16;
  const int nF = 12;
  double *base, *out;
  cudaMalloc(&base, sizeof(double) * nF * b.fs);
  cudaMalloc(&out, sizeof(double) * 8192);
  std::vector<double> h(nF * b.fs);
  for (size_t q = 0; q < h.size(); ++q) h[q] = (double)q;
  cudaMemcpy(base, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(Probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8 + 64);
  {
    const int s[3] = {bx, by, nf};
    CUtensorMap m;
    std::string err;
    if (EncodeBlockMap(&m, b, base, nF, s[0], s[1], s[2], &err)) { printf("encode %d %d %d: %s\n", s[0], s[1], s[2], err.c_str()); return 0; }
    const int n = s[0] * s[1] * s[2];
    Probe<<<1, 128, 8192 * 8 + 64>>>(m, c0, 1, 3, 2, n, out);
    cudaError_t e = cudaDeviceSynchronize();
    double v[2] = {0, 0};
    if (e == cudaSuccess) cudaMemcpy(v, out, 16, cudaMemcpyDeviceToHost);
    const double expect = 2.0 * b.fs + 3.0 * b.sk + 1.0 * b.sj + c0;
    printf("box %d x %d x 1 x %d c0=%d: %s  first=%.0f expect=%.0f\n", s[0], s[1], s[2], c0, cudaGetErrorString(e), v[0], expect);
  }
  return 0;
}
