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
  unsigned long long *table;     // multimap slots: tag32 | id32, 4-slot buckets
  uint64_t n_buckets;
  uint2 *cands;                  // (heavy, light)
  unsigned long long *cand_count;
  uint64_t cand_cap;
  uint32_t *graft_cand;
  unsigned long long *fstats;    // [0] entries stored [1] lookups [2] candidates [3] verified (ed<=2)
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

__global__ void k_fj_flags(const uint32_t *label, const unsigned long long *mass, uint64_t boundary, uint32_t n, uint8_t *is_light,
                           uint32_t *graft_cand, uint32_t *counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < n;
  const bool light = in && mass[label[i]] < boundary;
  if (in) { is_light[i] = light ? 1 : 0; graft_cand[i] = kNone; }
  const uint32_t ml = __ballot_sync(kFull, light), mh = __ballot_sync(kFull, in && !light);
  if ((threadIdx.x & 31u) == 0) {
    if (ml) atomicAdd(&counts[0], __popc(ml));
    if (mh) atomicAdd(&counts[1], __popc(mh));
  }
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
    const unsigned long long val = (h << 32) | a;                      // tag = low 32 bits of the hash
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
        tag = static_cast<uint32_t>(h);
        b = __umul64hi(h, J.n_buckets);
        lookups++;
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
          else if (static_cast<uint32_t>(sv[s] >> 32) == tag) {
            const uint32_t l = static_cast<uint32_t>(sv[s]);
            const uint32_t Ll = J.len[l];
            const uint32_t dl = Ll > L ? Ll - L : L - Ll;
            if (dl <= 2 && J.graft_cand[l] > a) {
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

// heavy pass, step 2: exact decision ed(h, l) <= 2 with a banded (±2) unit-cost DP, one pair per thread.
// rows = positions of h, band index k <-> column c = r - 2 + k of l.
__global__ void __launch_bounds__(256) k_fj_verify(JoinParams J, uint64_t m) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint2 pr = J.cands[i];
  const uint32_t a = pr.x, l = pr.y;
  if (J.graft_cand[l] <= a) return;                                    // already grafted on an earlier heavy amplicon
  const uint64_t *hw = J.words + static_cast<uint64_t>(a) * J.stride;
  const uint64_t *lw = J.words + static_cast<uint64_t>(l) * J.stride;
  const int Lh = static_cast<int>(J.len[a]), Ll = static_cast<int>(J.len[l]);
  // D[-1][c] = c + 1 (row "-1" is the empty prefix of h): band slots for row 0 hold columns -3..2 of row -1
  int prev[6];                                                         // prev[k] = D[r-1][c = (r-1) - 2 + k], k = 0..5
#pragma unroll
  for (int k = 0; k < 6; ++k) { const int c = -3 + k; prev[k] = (c >= -1 && c < Ll) ? c + 1 : 99; }
  // window of l's bases: bits 2k = base at column r - 2 + k
  uint32_t lwin = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) { const int c = -2 + k; if (c >= 0 && c < Ll) lwin |= base_at(lw, static_cast<uint32_t>(c)) << (2 * k); }
  uint64_t hword = 0;
  bool ok = true;
  for (int r = 0; r < Lh; ++r) {
    if ((r & 31) == 0) hword = hw[r >> 5];
    const uint32_t hb = static_cast<uint32_t>(hword >> ((r & 31) << 1)) & 3u;
    int rowmin = 99;
    int cur[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int c = r - 2 + k;
      int v = 99;
      if (c >= 0 && c < Ll) {
        const int diag = (c == 0) ? r : prev[k];                       // D[r-1][c-1]; for c == 0 it is D[r-1][-1] = r
        const int up = prev[k + 1];                                    // D[r-1][c]
        const int lf = (c == 0) ? r + 1 : ((k == 0) ? 99 : cur[k - 1]); // D[r][c-1]; for c == 0 it is D[r][-1] = r + 1
        const uint32_t lb = (lwin >> (2 * k)) & 3u;
        v = min(diag + (lb == hb ? 0 : 1), min(up, lf) + 1);
      }
      cur[k] = v;
      rowmin = min(rowmin, v);
    }
    if (rowmin > 2) { ok = false; break; }
#pragma unroll
    for (int k = 0; k < 5; ++k) prev[k] = cur[k];
    prev[5] = 99;                                                      // D[r][r+3] is outside the band
    lwin >>= 2;
    { const int c = r + 1 + 2; if (c < Ll) lwin |= base_at(lw, static_cast<uint32_t>(c)) << 8; }
  }
  if (!ok) return;
  // result = D[Lh-1][Ll-1], band index of column Ll-1 in row Lh-1 (prev[] holds that row now, shifted by one row)
  const int kend = (Ll - 1) - (Lh - 1) + 2;
  int dist = 99;
#pragma unroll
  for (int k = 0; k < 5; ++k) if (k == kend) dist = prev[k];
  if (dist <= 2) atomicMin(&J.graft_cand[l], a);
}

}  // namespace swb
