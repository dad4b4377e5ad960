// lat_probe.cu -- latency constants the LU-SGS wavefront design depends on, measured with clock64 in
// one thread block on an otherwise idle B200: dependent DFMA / MUFU.RCP64H chains, LDG from L1 / L2 /
// DRAM, __syncthreads with 6 warps, st.release.gpu with stores in flight, shared-memory round trip.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *buf, long long n, int *flag, long long *out) {
  const int tid = threadIdx.x;
  long long t0, t1;
  double a = buf[tid], b = 1.0000001, c = 1e-9;
  // 1. dependent DFMA chain
  __syncthreads();
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (tid == 0) out[0] = (t1 - t0);  // /256
  // 2. dependent rcp.approx + 2 newton
  __syncthreads();
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    y = fma(y, fma(-a, y, 1.0), y);
    y = fma(y, fma(-a, y, 1.0), y);
    a = y + 1.5;
  }
  t1 = clock64();
  if (tid == 0) out[1] = (t1 - t0);  // /64
  // 3. pointer chase through global memory: stride large -> DRAM/L2 miss first pass, L2 hit second, L1 third
  long long *p = reinterpret_cast<long long *>(buf) + n;  // chase array prepared by host
  long long idx = tid == 0 ? 0 : 0;
  __syncthreads();
  if (tid == 0) {
    t0 = clock64();
    for (int i = 0; i < 64; ++i) idx = __ldcg(p + idx);
    t1 = clock64();
    out[2] = (t1 - t0);  // cold: DRAM
    idx = 0;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) idx = __ldcg(p + idx);
    t1 = clock64();
    out[3] = (t1 - t0) + (idx == 12345);  // L2 hit (.cg)
    idx = 0;
    for (int i = 0; i < 64; ++i) idx = p[idx];
    idx = 0;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) idx = p[idx];
    t1 = clock64();
    out[4] = (t1 - t0) + (idx == 12345);  // L1 hit
  }
  // 4. __syncthreads x 64
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) out[5] = (t1 - t0);
  // 5. st.release.gpu after a burst of 5 st.cg per thread (as the publisher sees it)
  __syncthreads();
  long long acc = 0;
  for (int i = 0; i < 16; ++i) {
    for (int e = 0; e < 5; ++e) __stcg(buf + (e * 1024 + tid) + 8192 * (i & 1), a + e);
    __syncthreads();
    if (tid == 192) {
      t0 = clock64();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(i) : "memory");
      t1 = clock64();
      acc += t1 - t0;
    }
    __syncthreads();
  }
  if (tid == 192) out[6] = acc;  // /16
  // 6. shared-memory handoff: write, barrier, read, dependent, 64 times
  __shared__ double s[256];
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < 64; ++i) {
    s[tid] = a;
    __syncthreads();
    a = s[(tid + 1) & 191] + 1.0;
  }
  t1 = clock64();
  if (tid == 0) out[7] = (t1 - t0);
  // 7. ld.relaxed.gpu poll of a flag that is already set
  if (tid == 0) {
    t0 = clock64();
    int v = 0;
    for (int i = 0; i < 16; ++i) {
      asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag + (v & 0)) : "memory");
    }
    t1 = clock64();
    out[8] = (t1 - t0) + (v == 12345);
  }
  buf[tid] = a;
}
int main() {
  const long long n = 1 << 20;
  double *buf; int *flag; long long *out;
  cudaMalloc(&buf, sizeof(double) * n + sizeof(long long) * (64 * 4096 + 16));
  cudaMalloc(&flag, 64);
  cudaMallocManaged(&out, sizeof(long long) * 16);
  cudaMemset(buf, 0, sizeof(double) * n);
  // chase: element i*4096 -> (i+1)*4096 (32 KB apart: a new line and page region each hop)
  static long long h[64 * 4096 + 16];
  for (int i = 0; i < 64; ++i) h[i * 4096] = (long long)((i + 1) % 64) * 4096;
  cudaMemcpy(reinterpret_cast<long long *>(buf) + n, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<1, 224>>>(buf, n, flag, out);
    cudaDeviceSynchronize();
  }
  printf("DFMA dependent latency      %.1f cyc\n", out[0] / 256.0);
  printf("FastRcp + add dependent     %.1f cyc\n", out[1] / 64.0);
  printf("LDG chase first pass        %.1f cyc\n", out[2] / 64.0);
  printf("LDG.cg chase (L2 hit)       %.1f cyc\n", out[3] / 64.0);
  printf("LDG chase (L1 hit)          %.1f cyc\n", out[4] / 64.0);
  printf("__syncthreads (7 warps)     %.1f cyc\n", out[5] / 64.0);
  printf("st.release.gpu after stores %.1f cyc\n", out[6] / 16.0);
  printf("smem write+bar+read         %.1f cyc\n", out[7] / 64.0);
  printf("ld.relaxed.gpu flag         %.1f cyc\n", out[8] / 16.0);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
