// swarm_b200/csrc/common.cuh — shared device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace swb {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr unsigned kFull = 0xFFFFFFFFu;

// One 16-byte slot of the open-addressing amplicon table: a bucket step is ONE 128-bit load.
// Replaces the reference's SoA occupancy bitmap + hash_values[] + hash_data[]
// (/root/reference src/hashtable.cc:41-44,125-146).  id == kNone marks an empty slot.
struct __align__(16) Slot {
  uint64_t hash;
  uint32_t id;
  uint32_t len;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk; SASS: UBLKCP) --------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy, completion counted in bytes on `bar`. dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- cache-hinted loads ------------------------------------------------------------------------
// filter word: read-only, random, must stay L2-resident.  Measured on B200 (profiles/r1b_*): with
// `.L1::no_allocate` (SASS LDG.E.NA) the filter lines were NOT retained in L2 once bucket walks streamed
// through it (15.4 GB DRAM reads per 4 M seeds, L2 hit 61 %); the plain non-coherent load keeps them
// resident (2.6 GB, L2 hit 91.5 %).
__device__ __forceinline__ uint2 ld_filter(const uint2 *p) {
  uint2 v;
  asm("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_slot(const Slot *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
  lo = __shfl_sync(kFull, lo, src);
  hi = __shfl_sync(kFull, hi, src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
  uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
  lo = __shfl_up_sync(kFull, lo, d);
  hi = __shfl_up_sync(kFull, hi, d);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
  lo = __shfl_xor_sync(kFull, lo, m);
  hi = __shfl_xor_sync(kFull, hi, m);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// The filter pattern: 4 bits of one 64-bit block, two in each 32-bit half, chosen by the HIGH half of
// the hash (the block address uses the LOW half, so address and pattern are independent).
// Same structure as the reference's blocked filter (one 64-bit word per key, src/bloompat.cc:46-71),
// with the pattern computed instead of looked up in a 1024-entry table.
__device__ __forceinline__ uint2 filter_pattern(uint64_t h) {
  const uint32_t hi = static_cast<uint32_t>(h >> 32);
  uint2 m;
  m.x = (1u << (hi & 31u)) | (1u << ((hi >> 5) & 31u));
  m.y = (1u << ((hi >> 10) & 31u)) | (1u << ((hi >> 15) & 31u));
  return m;
}

// per-warp staging of output pairs in shared memory: one global atomicAdd per ~100 pairs instead of
// one per warp step (a single hot counter serialises at ~0.5 G atomics/s: 2.5 M atomics cost 5 ms)
constexpr int kStageCap = 160;
struct PairStage { uint2 buf[kStageCap]; };

__device__ __forceinline__ void stage_flush(PairStage &S, uint32_t &cnt, uint2 *out, unsigned long long *counter, uint64_t cap,
                                            uint32_t lane) {
  if (cnt == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(counter, static_cast<unsigned long long>(cnt));
  base = shfl_u64(base, 0);
  for (uint32_t i = lane; i < cnt; i += 32)
    if (base + i < cap) out[base + i] = S.buf[i];
  __syncwarp();
  cnt = 0;
}
// every lane contributes `mine` (0..2) pairs; warp-uniform bookkeeping, no global atomics
__device__ __forceinline__ void stage_push(PairStage &S, uint32_t &cnt, uint32_t mine, uint2 p0, uint2 p1, uint2 *out,
                                           unsigned long long *counter, uint64_t cap, uint32_t lane) {
  const uint32_t b1 = __ballot_sync(kFull, mine >= 1), b2 = __ballot_sync(kFull, mine >= 2);
  const uint32_t total = __popc(b1) + __popc(b2);
  if (total == 0) return;
  if (cnt + total > kStageCap) stage_flush(S, cnt, out, counter, cap, lane);
  const uint32_t lt = (1u << lane) - 1u;
  if (mine >= 1) S.buf[cnt + __popc(b1 & lt)] = p0;
  if (mine >= 2) S.buf[cnt + __popc(b1) + __popc(b2 & lt)] = p1;
  __syncwarp();
  cnt += total;
}

}  // namespace swb
