// tools/gather_bench.cu — microbenchmark: random 8-byte gathers over a buffer of S MiB (the filter
// access pattern of k_d1_network) with different load flavours.  Measures G loads/s on B200 to locate
// the L2-resident / DRAM knee and the L1tex gather ceiling.   nvcc -arch=sm_100a -O3 -o gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
template <int MODE>
__device__ __forceinline__ uint2 ld(const uint2 *p, uint64_t pol) {
  uint2 v;
  if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
  if (MODE == 4) asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  if (MODE == 5) asm volatile("ld.global.L1::evict_last.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
template <int MODE, int U>
__global__ void __launch_bounds__(256) k(const uint2 *buf, uint32_t mask, int iters, uint32_t *out) {
  uint64_t pol = 0;
  if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  uint64_t s = mix(blockIdx.x * 1315423911ull + threadIdx.x * 2654435761ull + 12345);
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      v[u] = ld<MODE>(buf + ((uint32_t)(s >> 33) & mask), pol);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y;
  }
  if (acc == 0x12345678) out[0] = acc;
}
template <int MODE, int U>
double run(const uint2 *buf, uint32_t mask, int grid, int iters, uint32_t *out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, U><<<grid, 256>>>(buf, mask, iters / 4, out);
  cudaEventRecord(a);
  k<MODE, U><<<grid, 256>>>(buf, mask, iters, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return (double)grid * 256 * iters * U / (ms * 1e-3) / 1e9;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int l2 = 0, pl2 = 0; cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0); cudaDeviceGetAttribute(&pl2, cudaDevAttrMaxPersistingL2CacheSize, 0);
  size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
  printf("%s SMs=%d L2=%d MB persistingL2max=%d MB fetchGranularity=%zu\n", pr.name, pr.multiProcessorCount, l2 >> 20, pl2 >> 20, lim);
  uint2 *buf; uint32_t *out; size_t maxb = 512ull << 20;
  cudaMalloc(&buf, maxb); cudaMemset(buf, 1, maxb); cudaMalloc(&out, 4);
  const int grids[] = {148 * 2, 148 * 4, 148 * 8};
  printf("%-8s %-6s", "MiB", "grid");
  const char *names[] = {"nc.NA", "plain", "cg", "nc.NA+evict_last", "nc", "L1evict_last"};
  for (auto n : names) printf(" %18s", n);
  printf("   (G loads/s, U=4)\n");
  for (int mb : {2, 4, 8, 16, 32, 48, 64, 96, 128, 256, 512}) {
    uint32_t mask = (uint32_t)(((size_t)mb << 20) / 8 - 1);
    // non power of two sizes: use largest pow2 below, fine for the knee
    uint32_t m2 = 1; while ((m2 << 1) <= mask + 1) m2 <<= 1; mask = m2 - 1;
    for (int g : grids) {
      printf("%-8d %-6d", (int)(((size_t)(mask + 1) * 8) >> 20), g);
      printf(" %18.1f", run<0, 4>(buf, mask, g, 2000, out));
      printf(" %18.1f", run<1, 4>(buf, mask, g, 2000, out));
      printf(" %18.1f", run<2, 4>(buf, mask, g, 2000, out));
      printf(" %18.1f", run<3, 4>(buf, mask, g, 2000, out));
      printf(" %18.1f", run<4, 4>(buf, mask, g, 2000, out));
      printf(" %18.1f", run<5, 4>(buf, mask, g, 2000, out));
      printf("\n");
    }
  }
  // unroll sweep at 16 MiB
  uint32_t mask = (16u << 20) / 8 - 1;
  printf("U sweep @16MiB grid=1184 nc.NA: U1 %.1f U2 %.1f U4 %.1f U8 %.1f\n", run<0, 1>(buf, mask, 1184, 2000, out), run<0, 2>(buf, mask, 1184, 2000, out),
         run<0, 4>(buf, mask, 1184, 2000, out), run<0, 8>(buf, mask, 1184, 2000, out));
  return 0;
}
