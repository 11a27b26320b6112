// swarm_b200/csrc/d1_network_v2.cuh — lean neighbour-network kernel for the HALF enumeration.
//
// Same result as k_d1_network<HALF> (d1_kernels.cuh) with ~2.5x fewer instructions per probe.  The
// first version was issue-bound (ncu, profiles/r1a: 82 warp-instructions per warp-wide filter load, 39 %
// issue-slot utilisation), while the memory system allows one L2-resident 8-byte gather per clock per
// SM (profiles/r1b_gather_microbench.txt).  What changed:
//   * a FUSED per-(position, base) table in shared memory, 32 bytes per entry, one address + two
//     LDS.128 per position instead of 5 scattered LDS.64 + selects:
//        D1 = Z[p][s]^Z[p][s+1]     (substitution arc s -> s+1)
//        D2 = Z[p][s]^Z[p][s+2]     (substitution arc s -> s+2, used when s < 2)
//        E  = Z[p][s]^Z[p-1][s]     (advances the running deletion hash)
//        Zm = Z[p-1][s]
//     with R(p) = PA(p)^TB^PB(p):  hash(sub) = H^D,  hash(del p) = R(p)^Zm,  R(p+1) = R(p)^E;
//   * one warp scan (exclusive XOR of E) instead of three, H read from the index kernel's output;
//   * the lane's bases live in one 64-bit register (shift by 2 per position);
//   * no ballots in the hot loop: a lane that sees a filter pass (~1 % of probes) appends to the warp's
//     shared-memory queue with a shared atomic; the queue level is checked once per position.
// Requires (longest+2)*128 bytes of shared memory for the table: used when that is <= 56 KB
// (sequences up to ~440 nt); longer inputs and the FULL enumeration use k_d1_network.
#pragma once
#include "d1_kernels.cuh"

namespace swb {

constexpr int kQueueCap2 = 192;   // survivors per warp: one position can add up to 3*32
constexpr uint32_t kTStride = 144; // bytes per position in the fused table: 128 + 16 padding, so that with an ODD number of
                                   // positions per lane the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups

struct WarpScratch2 {
  uint64_t qhash[kQueueCap2];
  uint32_t qcode[kQueueCap2];
  uint2 edges[kEdgeCap];
  uint32_t qn;
  uint32_t pad;
};

template <bool STATS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3) k_d1_network_half(D1Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [T: zlen positions x 144 B][per-warp 2 * batch*stride u64][per-warp WarpScratch2][per-warp 2 mbarriers]
  unsigned char *T = smem_raw;
  const uint32_t buf_words = P.batch * P.stride;
  uint64_t *bufs = reinterpret_cast<uint64_t *>(T + static_cast<size_t>(P.zlen) * kTStride);
  WarpScratch2 *scr = reinterpret_cast<WarpScratch2 *>(bufs + static_cast<size_t>(kWarpsPerCta) * 2 * buf_words);
  uint64_t *bars = reinterpret_cast<uint64_t *>(scr + kWarpsPerCta);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < P.zlen * 4; i += blockDim.x) {
    const uint32_t p = i >> 2, s = i & 3u;
    const uint64_t z = P.ztab[i];
    const uint64_t zm = p ? P.ztab[(p - 1) * 4 + s] : 0ull;
    ulonglong2 *e = reinterpret_cast<ulonglong2 *>(T + p * kTStride + s * 32);
    e[0] = make_ulonglong2(z ^ P.ztab[p * 4 + ((s + 1) & 3u)], z ^ P.ztab[p * 4 + ((s + 2) & 3u)]);
    e[1] = make_ulonglong2(z ^ zm, zm);
  }
  if (lane == 0) { mbar_init(&bars[warp * 2], 1); mbar_init(&bars[warp * 2 + 1], 1); scr[warp].qn = 0; }
  mbar_fence_init();
  __syncthreads();

  WarpScratch2 &S = scr[warp];
  uint64_t *mybuf = bufs + static_cast<size_t>(warp) * 2 * buf_words;
  uint64_t *mybar = &bars[warp * 2];
  const uint32_t G = gridDim.x * kWarpsPerCta;
  const uint32_t g = blockIdx.x * kWarpsPerCta + warp;
  const uint32_t first_batch = P.seed_begin / P.batch;
  const uint32_t n_batches = (P.seed_end - P.seed_begin + P.batch - 1) / P.batch;
  const uint32_t buf_bytes = buf_words * 8;
  const uint32_t t_base = smem_u32(T);

  unsigned long long st_var = 0, st_pass = 0, st_slots = 0, st_cmp = 0;
  uint32_t en = 0;
  uint32_t phase0 = 0, phase1 = 0;

  uint32_t bi = g;
  if (bi < n_batches && lane == 0) {
    mbar_expect_tx(&mybar[0], buf_bytes);
    tma_load_1d(mybuf, P.words + static_cast<uint64_t>(first_batch + bi) * buf_words, buf_bytes, &mybar[0]);
  }
  uint32_t cur = 0;
  for (; bi < n_batches; bi += G) {
    const uint32_t nb = bi + G;
    if (nb < n_batches && lane == 0) {
      mbar_expect_tx(&mybar[cur ^ 1], buf_bytes);
      tma_load_1d(mybuf + (cur ^ 1) * buf_words, P.words + static_cast<uint64_t>(first_batch + nb) * buf_words, buf_bytes,
                  &mybar[cur ^ 1]);
    }
    if (cur == 0) { mbar_wait(&mybar[0], phase0); phase0 ^= 1; } else { mbar_wait(&mybar[1], phase1); phase1 ^= 1; }
    const uint64_t *tile = mybuf + cur * buf_words;
    const uint32_t seed0 = (first_batch + bi) * P.batch;
    for (uint32_t k = 0; k < P.batch; ++k) {
      const uint32_t seed = seed0 + k;
      if (seed < P.seed_begin || seed >= P.seed_end) continue;
      const uint32_t L = P.len[seed];
      const uint64_t TA = P.hashes[seed];                  // H(seed), computed by k_d1_index
      const uint64_t *sw = tile + k * P.stride;
      const uint32_t c = ((L + 31) >> 5) | 1u;             // positions per lane, forced odd (bank spread), <= 33
      const uint32_t p0 = lane * c;
      // this lane's bases: chunk bits [2i, 2i+1] = base at p0+i; prev = base at p0-1 (4 = none)
      uint64_t chunk = 0;
      uint32_t prev = 4;
      if (p0 < L) {
        const uint32_t wi = p0 >> 5, sh = (p0 & 31u) << 1;
        const uint64_t w0 = sw[wi];
        const uint64_t w1 = (wi + 1 < P.stride) ? sw[wi + 1] : 0ull;
        chunk = sh ? ((w0 >> sh) | (w1 << (64 - sh))) : w0;
        if (p0) prev = base_at(sw, p0 - 1);
      }
      // exclusive XOR scan of E over positions -> R(p0) = TB ^ prefixE(p0), TB = TA ^ totalE
      uint64_t lE = 0;
      {
        uint64_t cb = chunk;
        for (uint32_t i = 0; i < c; ++i) {
          const uint32_t p = p0 + i;
          if (p < L) lE ^= reinterpret_cast<const ulonglong2 *>(T + p * kTStride + (static_cast<uint32_t>(cb) & 3u) * 32)[1].x;
          cb >>= 2;
        }
      }
      uint64_t iE = lE;
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        const uint64_t t = shfl_up_u64(iE, dd);
        if (lane >= static_cast<uint32_t>(dd)) iE ^= t;
      }
      const uint64_t TB = TA ^ shfl_u64(iE, 31);
      uint64_t R = TB ^ iE ^ lE;

      uint32_t taddr = t_base + p0 * kTStride;
      for (uint32_t it = 0; it < c; ++it, taddr += kTStride) {  // warp-uniform trip count
        const uint32_t p = p0 + it;
        const bool real = p < L;
        const uint32_t s = static_cast<uint32_t>(chunk) & 3u;
        chunk >>= 2;
        ulonglong2 d12, ezm;
        {
          const uint32_t a = taddr + s * 32;
          asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d12.x), "=l"(d12.y) : "r"(real ? a : t_base));
          asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(ezm.x), "=l"(ezm.y) : "r"(real ? a : t_base));
        }
        const uint64_t hA = TA ^ d12.x, hB = TA ^ d12.y, hC = R ^ ezm.y;
        const bool vA = real, vB = real && s < 2u, vC = real && (p == 0 || s != prev);
        uint2 wA = make_uint2(0, 0), wB = wA, wC = wA;
        if (vA) wA = ld_filter(P.filter + (static_cast<uint32_t>(hA) & P.filter_mask));
        if (vB) wB = ld_filter(P.filter + (static_cast<uint32_t>(hB) & P.filter_mask));
        if (vC) wC = ld_filter(P.filter + (static_cast<uint32_t>(hC) & P.filter_mask));
        const uint2 mA = filter_pattern(hA), mB = filter_pattern(hB), mC = filter_pattern(hC);
        const bool pA = vA && (((mA.x & ~wA.x) | (mA.y & ~wA.y)) == 0u);
        const bool pB = vB && (((mB.x & ~wB.x) | (mB.y & ~wB.y)) == 0u);
        const bool pC = vC && (((mC.x & ~wC.x) | (mC.y & ~wC.y)) == 0u);
        if (STATS) st_var += (vA ? 1u : 0u) + (vB ? 1u : 0u) + (vC ? 1u : 0u);
        if (pA || pB || pC) {                              // rare (~3 % of lanes): append to the warp queue
          const uint32_t cnt = (pA ? 1u : 0u) + (pB ? 1u : 0u) + (pC ? 1u : 0u);
          uint32_t at = atomicAdd(&S.qn, cnt);
          if (pA) { S.qhash[at] = hA; S.qcode[at] = (((s + 1u) & 3u) << 28) | p; ++at; }
          if (pB) { S.qhash[at] = hB; S.qcode[at] = (((s + 2u) & 3u) << 28) | p; ++at; }
          if (pC) { S.qhash[at] = hC; S.qcode[at] = (1u << 30) | p; }
          if (STATS) st_pass += cnt;
        }
        if (real) { R ^= ezm.x; prev = s; }
        __syncwarp();
        uint32_t qn = *reinterpret_cast<volatile uint32_t *>(&S.qn);
        if (qn > kQueueCap2 - 96) {
          drain_queue<1, STATS, WarpScratch2>(P, S, sw, seed, L, qn, en, lane, st_slots, st_cmp);
          if (lane == 0) S.qn = 0;
          __syncwarp();
        }
      }
      uint32_t qn = *reinterpret_cast<volatile uint32_t *>(&S.qn);
      if (qn) {
        drain_queue<1, STATS, WarpScratch2>(P, S, sw, seed, L, qn, en, lane, st_slots, st_cmp);
        if (lane == 0) S.qn = 0;
        __syncwarp();
      }
    }
    __syncwarp();
    cur ^= 1;
  }
  flush_edges(P, S, en, lane);
  if (STATS && P.stats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      st_var += __shfl_xor_sync(kFull, st_var, m);
      st_pass += __shfl_xor_sync(kFull, st_pass, m);
      st_slots += __shfl_xor_sync(kFull, st_slots, m);
      st_cmp += __shfl_xor_sync(kFull, st_cmp, m);
    }
    if (lane == 0) {
      atomicAdd(&P.stats[0], st_var);
      atomicAdd(&P.stats[1], st_pass);
      atomicAdd(&P.stats[2], st_slots);
      atomicAdd(&P.stats[3], st_cmp);
    }
  }
}

}  // namespace swb
