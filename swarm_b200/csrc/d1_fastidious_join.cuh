// swarm_b200/csrc/d1_fastidious_join.cuh — the --fastidious graft search as a pigeonhole join.
//
// The reference decides "light amplicon l can be grafted on heavy amplicon h" by asking whether they
// share a microvariant, V(h) ∩ V(l) != ∅ (src/algod1.cc:374-552).  V(x) is the set
// of all sequences at Levenshtein distance exactly 1 from x (src/variants.cc:184-249 enumerates each of
// them once), so for two DISTINCT sequences
//        V(h) ∩ V(l) != ∅   <=>   1 <= ed(h, l) <= 2
// (ed = 2: the middle of any 2-edit script is a common microvariant; ed = 1: substitute a third base
// at the differing position, or substitute the deleted base before deleting it).  Hence
//        graft_cand[l] = min { h in a heavy swarm : ed(h, l) <= 2 }.
// Two edits cannot touch three disjoint pieces of l, so at least one of
//        P0 = l[0,K)      P1 = l[K,2K)      P2 = l[Ll-K, Ll)          (K = min(64, minlen/3))
// occurs verbatim in h: P0 at offset 0, P2 at the end, P1 at offset K+δ, |δ| <= 2.  The light pass
// stores the three K-mer hashes of every light amplicon in a small multimap (3 entries per light
// amplicon instead of ~7L microvariants); the heavy pass makes 7 lookups per heavy amplicon instead of
// ~7L probes, and every candidate pair is decided exactly by a banded (±2) unit-cost DP.
// Heavy amplicons are processed in ascending id chunks so that `graft_cand[l] <= h` prunes most later
// candidates.  The enumeration kernels of d1_fastidious.cuh remain the path for inputs whose shortest
// sequence is too short for three disjoint pieces, and are the cross-check in the tests.
#pragma once
#include "d1_fastidious.cuh"

namespace swb {

struct JoinParams {
  const uint64_t *words;
  const uint32_t *len;
  uint32_t n, stride;
  uint32_t K;
  const uint8_t *is_light;       // per amplicon
  unsigned long long *table;     // multimap slots: tag24 | len8 | id32 (len8 = low 8 bits of the light amplicon's length), 4-slot buckets
  uint64_t n_buckets;
  uint2 *cands;                  // (heavy, light)
  unsigned long long *cand_count;
  uint64_t cand_cap;
  uint32_t *graft_cand;
  unsigned long long *fstats;    // [0] entries stored [1] lookups [2] candidates [3] verified (ed<=2)
  uint32_t *overflow;            // set when a chunk produced more candidates than `cands` holds
  unsigned long long *bloom;     // blocked Bloom filter over the light pieces (L2-resident): one 64-bit word, 4 bits per piece
  uint64_t bloom_words;
};

// 2K bits starting at nucleotide `off`, hashed together with the piece id
__device__ __forceinline__ uint64_t piece_hash(const uint64_t *w, uint32_t stride, uint32_t off, uint32_t K, uint32_t piece) {
  const uint32_t wi = off >> 5, sh = (off & 31u) << 1;
  uint64_t x[2] = {0, 0};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint32_t need = K > 32u * j ? min(32u, K - 32u * j) : 0u;      // nucleotides in this 64-bit piece word
    if (need) {
      const uint64_t a = (wi + j < stride) ? w[wi + j] : 0ull;
      const uint64_t b = (sh && wi + j + 1 < stride) ? w[wi + j + 1] : 0ull;
      uint64_t v = sh ? ((a >> sh) | (b << (64 - sh))) : a;
      if (need < 32) v &= (1ull << (2 * need)) - 1ull;
      x[j] = v;
    }
  }
  uint64_t h = x[0] + 0x9E3779B97F4A7C15ull * (piece + 1);
  h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
  h ^= x[1];
  h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
  return h ^ (h >> 31);
}

__global__ void __launch_bounds__(256) k_fj_flags(const uint32_t *label, const unsigned long long *mass, uint64_t boundary, uint32_t n, uint8_t *is_light,
                           uint32_t *graft_cand, uint32_t *counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < n;
  const bool light = in && mass[label[i]] < boundary;
  if (in) { is_light[i] = light ? 1 : 0; graft_cand[i] = kNone; }
  // one atomic per CTA and counter (per warp they were 6 * 10^5 atomics on two addresses: 0.31 ms at 10 M, profiles/r2y_launches_c3.txt)
  const uint32_t nl = __syncthreads_count(light), nh = __syncthreads_count(in && !light);
  if (threadIdx.x == 0) {
    if (nl) atomicAdd(&counts[0], nl);
    if (nh) atomicAdd(&counts[1], nh);
  }
}

// Bloom word and pattern of a piece hash.  80 % of the heavy lookups find nothing; the filter (16 bits per light piece, ~11 MB at
// 10 M amplicons: it stays in L2) answers those without touching the 114 MB multimap (ncu, profiles/r2s_k_fj_candidates: the
// candidate kernel was bound by the latency of its dependent random reads, 40 % issue slots, 13 cycles of long-scoreboard stall
// per issue)
__device__ __forceinline__ uint64_t fj_bloom_word(const JoinParams &J, uint64_t h) { return __umul64hi(h * 0x9E3779B97F4A7C15ull, J.bloom_words); }
__device__ __forceinline__ unsigned long long fj_bloom_pattern(uint64_t h) {
  return (1ull << ((h >> 8) & 63u)) | (1ull << ((h >> 14) & 63u)) | (1ull << ((h >> 20) & 63u)) | (1ull << ((h >> 26) & 63u));
}

// light pass: three K-mer entries per light amplicon
__global__ void __launch_bounds__(256) k_fj_insert(JoinParams J) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= J.n || !J.is_light[a]) return;
  const uint64_t *w = J.words + static_cast<uint64_t>(a) * J.stride;
  const uint32_t L = J.len[a];
  const uint32_t offs[3] = {0u, J.K, L - J.K};
#pragma unroll
  for (uint32_t piece = 0; piece < 3; ++piece) {
    const uint64_t h = piece_hash(w, J.stride, offs[piece], J.K, piece);
    // tag = low 24 bits of the hash; the low 8 bits of the length ride along so that the heavy pass can apply the length filter
    // |Ll - Lh| <= 2 without a dependent random read of len[l] (15 % of its stall samples, profiles/r2y_k_fj_candidates.txt)
    const unsigned long long val = ((h & 0xFFFFFFull) << 40) | (static_cast<unsigned long long>(L & 0xFFu) << 32) | a;
    atomicOr(&J.bloom[fj_bloom_word(J, h)], fj_bloom_pattern(h));
    uint64_t b = __umul64hi(h, J.n_buckets);
    for (bool placed = false; !placed;) {
      unsigned long long *slot = J.table + b * 4;
#pragma unroll
      for (int s = 0; s < 4 && !placed; ++s)
        if (slot[s] == kT2Empty && atomicCAS(&slot[s], kT2Empty, val) == kT2Empty) placed = true;
      if (++b == J.n_buckets) b = 0;
    }
  }
  if (J.fstats) atomicAdd(&J.fstats[0], 3ull);
}

// heavy pass, step 1: 7 lookups per heavy amplicon of [a_begin, a_end) -> candidate (heavy, light) pairs.
// Warp-synchronous bucket walk; candidates are staged per warp in shared memory (one global atomic per
// ~100 pairs).
__global__ void __launch_bounds__(256) k_fj_candidates(JoinParams J, uint32_t a_begin, uint32_t a_end) {
  __shared__ PairStage stage[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  PairStage &S = stage[warp];
  uint32_t scnt = 0;
  const uint64_t total = static_cast<uint64_t>(a_end - a_begin) * 7;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t rounds = (total + nthreads - 1) / nthreads;
  unsigned long long lookups = 0, found = 0;
  const uint32_t K = J.K;
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t t = r * nthreads + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t a = a_begin + static_cast<uint32_t>(t / 7);
    const uint32_t q = static_cast<uint32_t>(t % 7);
    bool walking = t < total && !J.is_light[a];
    uint32_t L = 0, tag = 0;
    uint64_t b = 0;
    if (walking) {
      const uint64_t *w = J.words + static_cast<uint64_t>(a) * J.stride;
      L = J.len[a];
      uint32_t piece = 0, off = 0;
      if (q == 0) { piece = 0; off = 0; }
      else if (q == 1) { piece = 2; off = L - K; }
      else {
        piece = 1;
        const int o = static_cast<int>(K) + static_cast<int>(q) - 4;       // K-2 .. K+2
        if (o < 0 || static_cast<uint32_t>(o) + K > L) walking = false;
        off = static_cast<uint32_t>(o < 0 ? 0 : o);
      }
      if (walking) {
        const uint64_t h = piece_hash(w, J.stride, off, K, piece);
        tag = static_cast<uint32_t>(h) & 0xFFFFFFu;
        b = __umul64hi(h, J.n_buckets);
        lookups++;
        const unsigned long long pat = fj_bloom_pattern(h);
        if ((J.bloom[fj_bloom_word(J, h)] & pat) != pat) walking = false;        // no light amplicon has this piece
      }
    }
    while (__any_sync(kFull, walking)) {
      unsigned long long sv[4] = {kT2Empty, kT2Empty, kT2Empty, kT2Empty};
      if (walking) {
        const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(J.table + b * 4);
        const ulonglong2 x = bp[0], y = bp[1];
        sv[0] = x.x; sv[1] = x.y; sv[2] = y.x; sv[3] = y.y;
      }
      bool full = walking;
      uint32_t cv[4] = {0, 0, 0, 0};
      uint32_t nc = 0;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (full) {
          if (sv[s] == kT2Empty) full = false;
          else if (static_cast<uint32_t>(sv[s] >> 40) == tag) {
            const uint32_t l = static_cast<uint32_t>(sv[s]);
            const uint32_t dl = (static_cast<uint32_t>(sv[s] >> 32) - L) & 0xFFu;      // (Ll - Lh) mod 256: -2 .. 2 pass (k_fj_verify re-checks the true lengths)
            if ((dl <= 2u || dl >= 254u) && J.graft_cand[l] > a) {
              if (nc == 0) cv[0] = l; else if (nc == 1) cv[1] = l; else if (nc == 2) cv[2] = l; else cv[3] = l;
              ++nc;
            }
          }
        }
      }
      found += nc;
      stage_push(S, scnt, min(nc, 2u), make_uint2(a, cv[0]), make_uint2(a, cv[1]), J.cands, J.cand_count, J.cand_cap, lane);
      if (__any_sync(kFull, nc > 2))
        stage_push(S, scnt, nc > 2 ? nc - 2 : 0u, make_uint2(a, cv[2]), make_uint2(a, cv[3]), J.cands, J.cand_count, J.cand_cap, lane);
      if (walking) {
        if (!full) walking = false;
        else if (++b == J.n_buckets) b = 0;
      }
    }
  }
  stage_flush(S, scnt, J.cands, J.cand_count, J.cand_cap, lane);
  if (J.fstats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { lookups += __shfl_xor_sync(kFull, lookups, m); found += __shfl_xor_sync(kFull, found, m); }
    if (lane == 0) { atomicAdd(&J.fstats[1], lookups); if (found) atomicAdd(&J.fstats[2], found); }
  }
}

// 32 nucleotides of the packed sequence w starting at position p (zero beyond its words)
__device__ __forceinline__ uint64_t fj_window(const uint64_t *w, uint32_t stride, uint32_t p) {
  const uint32_t wi = p >> 5, sh = (p & 31u) << 1;
  const uint64_t a = wi < stride ? w[wi] : 0ull;
  if (sh == 0) return a;
  const uint64_t b = wi + 1 < stride ? w[wi + 1] : 0ull;
  return (a >> sh) | (b << (64 - sh));
}
// longest common extension: how many nucleotides of h[i..) and l[j..) agree, at most cap (32 per step: one XOR of two windows)
__device__ __forceinline__ int fj_lce(const uint64_t *h, const uint64_t *l, uint32_t stride, int i, int j, int cap) {
  int n = 0;
  while (n < cap) {
    const uint64_t x = fj_window(h, stride, static_cast<uint32_t>(i + n)) ^ fj_window(l, stride, static_cast<uint32_t>(j + n));
    if (x) { n += (__ffsll(static_cast<long long>(x)) - 1) >> 1; break; }
    n += 32;
  }
  return min(n, cap);
}

// heavy pass, step 2: exact decision ed(h, l) <= 2, one pair per thread, grid-stride over the candidates counted ON THE DEVICE
// (no host round trip between the two steps of a chunk).  r1 filled a banded (+-2) unit-cost DP row by row (~9 000
// instructions a pair, 8.4 of the 10.5 ms of the graft search at 10 M amplicons, profiles/r2r_k_fj_verify); this is the
// diagonal-wise formulation of the same distance (Landau-Vishkin): fr_e[d] = the furthest row of h reachable on diagonal d
// (column - row) with e edits; one edit moves to a neighbouring diagonal or one step down, then the path slides along the
// diagonal for as long as the sequences agree — a longest-common-extension query on the 2-bit packed words.  At most
// 1 + 3 + 5 diagonal states for e = 0, 1, 2; accepted iff the end (Lh, Ll) is reached.
__global__ void __launch_bounds__(256) k_fj_verify(JoinParams J) {
  const uint64_t m = min(*J.cand_count, static_cast<unsigned long long>(J.cand_cap));
  if (*J.cand_count > J.cand_cap && blockIdx.x == 0 && threadIdx.x == 0) *J.overflow = 1u;
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  unsigned long long verified = 0;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += nth) {
    const uint2 pr = J.cands[i];
    const uint32_t a = pr.x, l = pr.y;
    if (J.graft_cand[l] <= a) continue;                                  // already grafted on an earlier heavy amplicon
    const uint64_t *hw = J.words + static_cast<uint64_t>(a) * J.stride;
    const uint64_t *lw = J.words + static_cast<uint64_t>(l) * J.stride;
    const int Lh = static_cast<int>(J.len[a]), Ll = static_cast<int>(J.len[l]);
    const int dt = Ll - Lh;                                              // the diagonal of the end point
    if (dt < -2 || dt > 2) continue;                                     // (the candidates' length filter works modulo 256)
    constexpr int NONE = -100000;
    int fr[5] = {NONE, NONE, NONE, NONE, NONE};                          // index d + 2
    fr[2] = fj_lce(hw, lw, J.stride, 0, 0, min(Lh, Ll));
    bool ok = dt == 0 && fr[2] >= Lh;
    for (int e = 1; e <= 2 && !ok; ++e) {
      int nf[5] = {NONE, NONE, NONE, NONE, NONE};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int d = k - 2;
        if (d < -e || d > e) continue;
        int best = fr[k] != NONE ? fr[k] + 1 : NONE;                     // substitution: one step down the same diagonal
        if (k > 0 && fr[k - 1] != NONE) best = max(best, fr[k - 1]);     // one more base of l: arrive from diagonal d - 1, same row
        if (k < 4 && fr[k + 1] != NONE) best = max(best, fr[k + 1] + 1); // one more base of h: arrive from diagonal d + 1, next row
        if (best == NONE) continue;
        int r = min(best, min(Lh, Ll - d));
        if (r < 0 || r + d < 0) continue;
        r += fj_lce(hw, lw, J.stride, r, r + d, min(Lh - r, Ll - (r + d)));
        nf[k] = r;
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) fr[k] = nf[k];
      ok = fr[dt + 2] >= Lh;
    }
    if (ok) { atomicMin(&J.graft_cand[l], a); ++verified; }
  }
  if (J.fstats) {
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) verified += __shfl_xor_sync(kFull, verified, mm);
    if ((threadIdx.x & 31u) == 0 && verified) atomicAdd(&J.fstats[3], verified);
  }
}

}  // namespace swb
