// swarm_b200/csrc/d1_fastidious.cuh — the --fastidious graft search on the device.
//
// Reference (/root/reference src/algod1.cc:1291-1475): light swarms = mass < boundary.  Every
// microvariant of every light amplicon is entered into a huge Bloom filter (mark_light_var :495-518);
// every microvariant w of every amplicon h of a heavy swarm is tested against it (check_heavy_var
// :398-450) and, on a hit, w's own ~7L microvariants are generated and looked up in a table of the
// light amplicons (check_heavy_var_2 :374-395, hash_check_attach :339-371) to find WHICH light
// amplicon l shares w; graft_cand[l] = min h (add_graft_candidate :244-258).  The Bloom filter and
// the second-generation enumeration exist only because a CPU cannot afford to remember which light
// amplicon produced each variant.  With HBM we can: the result is exactly
//        graft_cand[l] = min { h in heavy swarms : V(h) ∩ V(l) != ∅ }          
// so the light pass stores (variant hash tag, variant code, l) in an open-addressing multimap
// (8-byte slots, 32-byte buckets = one DRAM sector), and the heavy pass probes it once per variant
// and verifies a tag match exactly by comparing the two virtual sequences h∘edit_h and l∘edit_l word
// by word.  No second generation, no false-positive blow-up.
#pragma once
#include "d1_kernels.cuh"

namespace swb {

constexpr unsigned long long kT2Empty = 0xFFFFFFFFFFFFFFFFull;

struct FastParams {
  D1Params P;
  const uint32_t *label;        // swarm seed per amplicon
  unsigned long long *mass;     // per amplicon id (only entries at seeds are meaningful)
  uint64_t boundary;
  uint32_t *light_ids, *heavy_ids;
  uint32_t *counts;             // [0] light, [1] heavy
  const uint32_t *ids;          // the list this launch works on
  uint32_t n_ids;
  unsigned long long *t2;       // multimap slots
  uint64_t n_buckets;           // 4 slots each
  uint32_t *graft_cand;
  unsigned long long *fstats;   // [0] light variants stored, [1] heavy variants probed, [2] tag matches, [3] verified
};

// slot = tag(15) | type(2) base(2) pos(13) | id(32)
__device__ __forceinline__ unsigned long long t2_pack(uint64_t h, uint32_t code, uint32_t id) {
  const uint32_t c17 = ((code >> 30) << 15) | (((code >> 28) & 3u) << 13) | (code & 0x1FFFu);
  return ((h >> 49) << 49) | (static_cast<unsigned long long>(c17) << 32) | id;
}
__device__ __forceinline__ uint64_t t2_bucket(uint64_t h, uint64_t n_buckets) {
  return __umul64hi(h << 15, n_buckets);      // bits 0..48 (the tag uses 49..63), uniform in [0, n_buckets)
}

__global__ void k_fast_mass(const uint32_t *label, const uint64_t *abundance, unsigned long long *mass, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&mass[label[i]], static_cast<unsigned long long>(abundance[i]));
}

// light / heavy id lists (src/algod1.cc:1307-1321 counts them; :466-491, :530-551 iterate them)
__global__ void k_fast_split(FastParams F) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const bool in = i < F.P.n;
  const bool light = in && F.mass[F.label[i]] < F.boundary;
  const bool heavy = in && !light;
  const uint32_t ml = __ballot_sync(kFull, light), mh = __ballot_sync(kFull, heavy);
  uint32_t bl = 0, bh = 0;
  if (lane == 0) {
    if (ml) bl = atomicAdd(&F.counts[0], __popc(ml));
    if (mh) bh = atomicAdd(&F.counts[1], __popc(mh));
  }
  bl = __shfl_sync(kFull, bl, 0);
  bh = __shfl_sync(kFull, bh, 0);
  if (light) F.light_ids[bl + __popc(ml & ((1u << lane) - 1u))] = i;
  if (heavy) F.heavy_ids[bh + __popc(mh & ((1u << lane) - 1u))] = i;
  if (in) F.graft_cand[i] = kNone;
}

// light pass: mark_light_var (src/algod1.cc:495-518) — store every microvariant of every light amplicon
__global__ void __launch_bounds__(256) k_fast_light(FastParams F) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *zs = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *seeds = zs + F.P.zlen * 4;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < F.P.zlen * 4; i += blockDim.x) zs[i] = F.P.ztab[i];
  __syncthreads();
  uint64_t *sw = seeds + warp * F.P.stride;
  const uint32_t G = gridDim.x * (blockDim.x >> 5);
  unsigned long long stored = 0;
  for (uint32_t k = blockIdx.x * (blockDim.x >> 5) + warp; k < F.n_ids; k += G) {
    const uint32_t a = F.ids[k];
    const uint32_t L = F.P.len[a];
    __syncwarp();
    for (uint32_t j = lane; j < F.P.stride; j += 32) sw[j] = F.P.words[static_cast<uint64_t>(a) * F.P.stride + j];
    __syncwarp();
    auto emit = [&](const bool (&vv)[9], const uint64_t (&vh)[9], const uint32_t (&vc)[9]) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        if (!vv[q]) continue;
        const unsigned long long val = t2_pack(vh[q], vc[q], a);
        uint64_t b = t2_bucket(vh[q], F.n_buckets);
        for (bool placed = false; !placed;) {
          unsigned long long *slot = F.t2 + b * 4;
#pragma unroll
          for (int s = 0; s < 4 && !placed; ++s)
            if (slot[s] == kT2Empty && atomicCAS(&slot[s], kT2Empty, val) == kT2Empty) placed = true;
          if (++b == F.n_buckets) b = 0;
        }
        stored++;
      }
    };
    enumerate_variants<0>(zs, sw, L, lane, emit);
  }
  if (F.fstats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) stored += __shfl_xor_sync(kFull, stored, m);
    if (lane == 0) atomicAdd(&F.fstats[0], stored);
  }
}

struct FastScratch {
  uint64_t qhash[kQueueCap];
  uint32_t qcode[kQueueCap];
  unsigned long long qval[kQueueCap];
};

// verify queued tag matches: is (heavy seed ∘ code_h) the same sequence as (light amplicon ∘ code_l)?
__device__ __forceinline__ void fast_drain(const FastParams &F, FastScratch &S, const uint64_t *sw, uint32_t seed, uint32_t L,
                                           uint32_t &qn, uint32_t lane, unsigned long long &verified) {
  const uint32_t sub = lane >> 3, j = lane & 7u;
  const uint32_t nw = (L + 31) >> 5;
  __syncwarp();
  for (uint32_t b0 = 0; b0 < qn; b0 += 4) {
    const uint32_t e = b0 + sub;
    const bool act = e < qn;
    const uint32_t code = act ? S.qcode[e] : 0u;
    const unsigned long long val = act ? S.qval[e] : 0ull;
    const uint32_t ht = code >> 30, hb = (code >> 28) & 3u, hp = code & 0x0FFFFFFFu;
    const uint32_t lid = static_cast<uint32_t>(val);
    const uint32_t c17 = static_cast<uint32_t>(val >> 32) & 0x1FFFFu;
    const uint32_t lt = c17 >> 15, lb = (c17 >> 13) & 3u, lp = c17 & 0x1FFFu;
    bool bad = false;
    if (act) {
      const uint32_t LL = F.P.len[lid];
      const uint32_t hv = ht == 0 ? L : (ht == 1 ? L - 1 : L + 1);
      const uint32_t lv = lt == 0 ? LL : (lt == 1 ? LL - 1 : LL + 1);
      if (hv != lv) bad = true;
      else {
        const uint64_t *lw = F.P.words + static_cast<uint64_t>(lid) * F.P.stride;
        const uint32_t lnw = (LL + 31) >> 5, vw = (hv + 31) >> 5;
        for (uint32_t k = j; k < vw; k += 8)
          if (variant_word(sw, nw, ht, hp, hb, k) != variant_word(lw, lnw, lt, lp, lb, k)) bad = true;
      }
    }
    const uint32_t bb = (__ballot_sync(kFull, bad) >> (sub * 8)) & 0xFFu;
    if (act && bb == 0 && j == 0) {
      atomicMin(&F.graft_cand[lid], seed);       // add_graft_candidate: keep the smallest heavy id
      verified++;
    }
  }
  __syncwarp();
  qn = 0;
}

// heavy pass: check_heavy_var (src/algod1.cc:398-450) without the second generation
__global__ void __launch_bounds__(256) k_fast_heavy(FastParams F) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *zs = reinterpret_cast<uint64_t *>(smem_raw);
  uint64_t *seeds = zs + F.P.zlen * 4;
  FastScratch *scr = reinterpret_cast<FastScratch *>(seeds + static_cast<size_t>(blockDim.x >> 5) * F.P.stride);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < F.P.zlen * 4; i += blockDim.x) zs[i] = F.P.ztab[i];
  __syncthreads();
  uint64_t *sw = seeds + warp * F.P.stride;
  FastScratch &S = scr[warp];
  const uint32_t G = gridDim.x * (blockDim.x >> 5);
  unsigned long long probed = 0, tagm = 0, verified = 0;
  uint32_t qn = 0;
  for (uint32_t k = blockIdx.x * (blockDim.x >> 5) + warp; k < F.n_ids; k += G) {
    const uint32_t a = F.ids[k];
    const uint32_t L = F.P.len[a];
    __syncwarp();
    for (uint32_t j = lane; j < F.P.stride; j += 32) sw[j] = F.P.words[static_cast<uint64_t>(a) * F.P.stride + j];
    __syncwarp();
    auto emit = [&](const bool (&vv)[9], const uint64_t (&vh)[9], const uint32_t (&vc)[9]) {
      // first bucket of every variant: one 32-byte sector = two 128-bit loads; three variants in flight
#pragma unroll
      for (int q0 = 0; q0 < 9; q0 += 3) {
        ulonglong2 lo[3], hi[3];
        uint64_t bk[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int q = q0 + r;
          bk[r] = t2_bucket(vh[q], F.n_buckets);
          lo[r] = make_ulonglong2(kT2Empty, kT2Empty);
          hi[r] = lo[r];
          if (vv[q]) {
            const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(F.t2 + bk[r] * 4);
            lo[r] = bp[0];
            hi[r] = bp[1];
          }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int q = q0 + r;
          if (vv[q]) probed++;
          const unsigned long long tagbits = (vh[q] >> 49) << 49;
          bool pending = vv[q];
          unsigned long long s0 = lo[r].x, s1 = lo[r].y, s2 = hi[r].x, s3 = hi[r].y;
          uint64_t b = bk[r];
          // warp-uniform loop over (rare) overflow buckets
          while (__any_sync(kFull, pending)) {
            const unsigned long long sv[4] = {s0, s1, s2, s3};
            bool full = true;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const bool m = pending && full && sv[s] != kT2Empty && ((sv[s] >> 49) << 49) == tagbits;
              const uint32_t bal = __ballot_sync(kFull, m);
              if (bal) {
                const uint32_t cnt = __popc(bal);
                if (qn + cnt > kQueueCap) fast_drain(F, S, sw, a, L, qn, lane, verified);
                if (m) {
                  const uint32_t at = qn + __popc(bal & ((1u << lane) - 1u));
                  S.qhash[at] = vh[q]; S.qcode[at] = vc[q]; S.qval[at] = sv[s];
                  tagm++;
                }
                qn += cnt;
              }
              if (sv[s] == kT2Empty) full = false;
            }
            if (pending) {
              if (!full) pending = false;
              else {
                if (++b == F.n_buckets) b = 0;
                const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(F.t2 + b * 4);
                const ulonglong2 x = bp[0], y = bp[1];
                s0 = x.x; s1 = x.y; s2 = y.x; s3 = y.y;
              }
            }
          }
        }
      }
    };
    enumerate_variants<0>(zs, sw, L, lane, emit);
    if (qn) fast_drain(F, S, sw, a, L, qn, lane, verified);
  }
  if (F.fstats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      probed += __shfl_xor_sync(kFull, probed, m);
      tagm += __shfl_xor_sync(kFull, tagm, m);
      verified += __shfl_xor_sync(kFull, verified, m);
    }
    if (lane == 0) {
      atomicAdd(&F.fstats[1], probed);
      atomicAdd(&F.fstats[2], tagm);
      atomicAdd(&F.fstats[3], verified);
    }
  }
}

}  // namespace swb
