// swarm_b200/csrc/d1_join.cuh — the d=1 neighbour network as a pigeonhole join (enum_mode JOIN).
//
// v is a microvariant of u  <=>  ed(u, v) = 1 (src/variants.cc:184-249 enumerates exactly the
// sequences at Levenshtein distance 1).  One edit cannot touch both of two disjoint pieces, so u and v
// share either their first K or their last K nucleotides verbatim (K = min(64, minlen/2); the prefix
// is anchored at the start, the suffix at the end, so indels do not shift them).  Instead of probing
// ~340 (HALF) or ~1017 (FULL) microvariant hashes per amplicon, every amplicon puts TWO K-mer entries
// into a multimap and makes TWO lookups; each candidate that shares a piece is decided exactly on the
// packed words (Hamming distance 1, or one shifted comparison for a length difference of 1).
// Result = the same directed link set as k_d1_network (tested).  Identical sequences (the reference's
// fatal duplicate check, src/algod1.cc:1141-1150) show up as ed = 0 candidates.
#pragma once
#include "d1_fastidious_join.cuh"

namespace swb {

struct NetJoinParams {
  const uint64_t *words;
  const uint32_t *len;
  const uint64_t *abundance;
  uint32_t n, stride, K;
  unsigned long long *table;       // tag32 | id32, 4-slot buckets; this rank holds buckets [b_lo, b_hi) at table[(b-b_lo)*4]
  uint64_t n_buckets;              // global bucket count (the hash -> bucket map is the same on every rank)
  uint64_t b_lo, b_hi;             // multi-GPU: the K-mer table is sharded by bucket range (hash range)
  uint2 *edges;
  unsigned long long *edge_count;
  uint64_t edge_cap;
  uint32_t seed_begin, seed_end;
  int ncb;
  uint32_t *dup_flag;
  uint2 *cands;
  unsigned long long *cand_count;
  uint64_t cand_cap;
  unsigned long long *stats;       // [0] lookups [1] candidates [2] slots visited [3] exact comparisons
};

__global__ void __launch_bounds__(256) k_join_index(NetJoinParams J) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t a = static_cast<uint32_t>(t >> 1), piece = static_cast<uint32_t>(t & 1u);
  if (a >= J.n) return;
  const uint64_t *w = J.words + static_cast<uint64_t>(a) * J.stride;
  const uint32_t L = J.len[a];
  const uint64_t h = piece_hash(w, J.stride, piece ? L - J.K : 0u, J.K, piece);
  const unsigned long long val = (h << 32) | a;
  uint64_t b = __umul64hi(h, J.n_buckets);
  if (b < J.b_lo || b >= J.b_hi) return;            // another rank owns this hash range
  for (;;) {
    unsigned long long *slot = J.table + (b - J.b_lo) * 4;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (slot[s] == kT2Empty && atomicCAS(&slot[s], kT2Empty, val) == kT2Empty) return;
    if (++b == J.b_hi) b = J.b_lo;                  // linear probing wraps inside the shard
  }
}

// exact test on packed words: 0 = identical, 1 = exactly one edit apart, 2 = further
__device__ __forceinline__ int edit_class(const uint64_t *x, uint32_t Lx, const uint64_t *y, uint32_t Ly, uint32_t stride) {
  if (Lx == Ly) {                                   // substitution: exactly one differing 2-bit group
    uint32_t diff = 0;
    const uint32_t nw = (Lx + 31) >> 5;
    for (uint32_t j = 0; j < nw; ++j) {
      const uint64_t d = x[j] ^ y[j];
      const uint64_t g = (d | (d >> 1)) & 0x5555555555555555ull;
      diff += __popcll(g);
      if (diff > 1) return 2;
    }
    return static_cast<int>(diff);
  }
  // one indel: make x the longer one; x[0,p) == y[0,p) and x[p+1, Lx) == y[p, Ly) for the first mismatch p
  if (Lx < Ly) { const uint64_t *t = x; x = y; y = t; const uint32_t tl = Lx; Lx = Ly; Ly = tl; }
  if (Lx != Ly + 1) return 2;
  const uint32_t nw = (Lx + 31) >> 5;
  uint32_t j = 0;
  while (j < nw && x[j] == y[j]) ++j;               // y is zero padded, so a clean prefix match can run to the end
  if (j == nw) return 1;                            // only the last base of x is extra (deleted at the end)
  // first differing group in word j
  const uint64_t d = x[j] ^ y[j];
  const uint32_t bit = static_cast<uint32_t>(__ffsll(static_cast<long long>((d | (d >> 1)) & 0x5555555555555555ull)) - 1);   // even bit index
  const uint32_t p = (j << 5) + (bit >> 1);
  if (p >= Lx) return 2;
  // compare x shifted down by one nucleotide from position p with y from position p
  for (uint32_t k = j; k < nw; ++k) {
    const uint64_t cur = x[k], nxt = (k + 1 < stride) ? x[k + 1] : 0ull;
    uint64_t xs;                                     // word k of del(x, p)
    if (k == j) {
      const uint32_t sh = bit;                       // bit offset of position p inside word j
      const uint64_t low = sh ? (cur & ((1ull << sh) - 1)) : 0ull;
      const uint64_t high = (sh < 62) ? ((cur >> (sh + 2)) << sh) : 0ull;
      xs = low | high | (nxt << 62);
    } else {
      xs = (cur >> 2) | (nxt << 62);
    }
    if (xs != y[k]) return 2;
  }
  return 1;
}

// step 1: two lookups per amplicon -> candidate pairs.  Only candidates with a larger id are kept
// (each unordered pair once, from the smaller id a).  The pair is stored as (a, v) when found through
// the prefix piece and as (v, a) when found through the suffix piece, so step 2 knows which lookup
// produced it.  No sequence is touched here: the walk is a short chain of 32-byte bucket reads.
__global__ void __launch_bounds__(256) k_join_candidates(NetJoinParams J) {
  __shared__ PairStage stage[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  PairStage &S = stage[warp];
  uint32_t scnt = 0;
  const uint64_t total_lookups = static_cast<uint64_t>(J.seed_end - J.seed_begin) * 2;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t rounds = (total_lookups + nthreads - 1) / nthreads;
  unsigned long long st_c = 0, st_s = 0, st_l = 0;
  const uint32_t K = J.K;
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t t = r * nthreads + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool in = t < total_lookups;
    const uint32_t a = J.seed_begin + static_cast<uint32_t>(t >> 1), piece = static_cast<uint32_t>(t & 1u);
    const uint64_t *w = J.words + static_cast<uint64_t>(in ? a : 0) * J.stride;
    const uint32_t L = in ? J.len[a] : 0u;
    const uint64_t h = in ? piece_hash(w, J.stride, piece ? L - K : 0u, K, piece) : 0ull;
    const uint32_t tag = static_cast<uint32_t>(h);
    uint64_t b = __umul64hi(h, J.n_buckets);
    bool walking = in && b >= J.b_lo && b < J.b_hi;   // only the owner of the hash range walks it
    if (walking) st_l++;
    while (__any_sync(kFull, walking)) {
      unsigned long long sv[4] = {kT2Empty, kT2Empty, kT2Empty, kT2Empty};
      if (walking) {
        const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(J.table + (b - J.b_lo) * 4);
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
          else {
            st_s++;
            const uint32_t v = static_cast<uint32_t>(sv[s]);
            if (static_cast<uint32_t>(sv[s] >> 32) == tag && v > a) {
              if (nc == 0) cv[0] = v; else if (nc == 1) cv[1] = v; else if (nc == 2) cv[2] = v; else cv[3] = v;
              ++nc;
            }
          }
        }
      }
      st_c += nc;
      const uint2 q0 = piece ? make_uint2(cv[0], a) : make_uint2(a, cv[0]);
      const uint2 q1 = piece ? make_uint2(cv[1], a) : make_uint2(a, cv[1]);
      stage_push(S, scnt, min(nc, 2u), q0, q1, J.cands, J.cand_count, J.cand_cap, lane);
      if (__any_sync(kFull, nc > 2)) {
        const uint2 q2 = piece ? make_uint2(cv[2], a) : make_uint2(a, cv[2]);
        const uint2 q3 = piece ? make_uint2(cv[3], a) : make_uint2(a, cv[3]);
        stage_push(S, scnt, nc > 2 ? nc - 2 : 0u, q2, q3, J.cands, J.cand_count, J.cand_cap, lane);
      }
      if (walking) {
        if (!full) walking = false;
        else if (++b == J.b_hi) b = J.b_lo;
      }
    }
  }
  stage_flush(S, scnt, J.cands, J.cand_count, J.cand_cap, lane);
  if (J.stats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      st_c += __shfl_xor_sync(kFull, st_c, m);
      st_s += __shfl_xor_sync(kFull, st_s, m);
      st_l += __shfl_xor_sync(kFull, st_l, m);
    }
    if (lane == 0) {
      atomicAdd(&J.stats[0], st_l);
      atomicAdd(&J.stats[1], st_c);
      atomicAdd(&J.stats[2], st_s);
    }
  }
}

// step 2: one candidate pair per thread, decided exactly on the packed words
__global__ void __launch_bounds__(256) k_join_verify(NetJoinParams J, uint64_t m) {
  __shared__ PairStage stage[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  PairStage &S = stage[warp];
  uint32_t scnt = 0;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t rounds = (m + nthreads - 1) / nthreads;
  unsigned long long st_x = 0;
  const uint32_t K = J.K;
  const uint64_t kmask1 = K >= 64 ? ~0ull : (K > 32 ? (1ull << (2 * (K - 32))) - 1 : 0ull);
  const uint64_t kmask0 = K >= 32 ? ~0ull : (1ull << (2 * K)) - 1;
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t i = r * nthreads + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t mine = 0;
    uint2 e1 = make_uint2(0, 0), e2 = make_uint2(0, 0);
    if (i < m) {
      const uint2 pr = J.cands[i];
      const bool by_suffix = pr.x > pr.y;
      const uint32_t a = by_suffix ? pr.y : pr.x, v = by_suffix ? pr.x : pr.y;
      const uint32_t L = J.len[a], Lv = J.len[v];
      const uint32_t dl = Lv > L ? Lv - L : L - Lv;
      if (dl <= 1) {
        const uint64_t *w = J.words + static_cast<uint64_t>(a) * J.stride;
        const uint64_t *vw = J.words + static_cast<uint64_t>(v) * J.stride;
        bool skip = false;
        if (by_suffix)                                     // also shares the prefix piece -> the prefix lookup owns the pair
          skip = ((w[0] ^ vw[0]) & kmask0) == 0 && (K <= 32 || ((w[1] ^ vw[1]) & kmask1) == 0);
        if (!skip) {
          st_x++;
          const int cls = edit_class(w, L, vw, Lv, J.stride);
          if (cls == 0) atomicExch(J.dup_flag, 1u);
          if (cls == 1) {
            const uint64_t aa = J.abundance[a], av = J.abundance[v];
            const bool f = J.ncb || aa >= av, g = J.ncb || av >= aa;
            if (f) { e1 = make_uint2(a, v); mine = 1; }
            if (g) { if (mine) e2 = make_uint2(v, a); else e1 = make_uint2(v, a); ++mine; }
          }
        }
      }
    }
    stage_push(S, scnt, mine, e1, e2, J.edges, J.edge_count, J.edge_cap, lane);
  }
  stage_flush(S, scnt, J.edges, J.edge_count, J.edge_cap, lane);
  if (J.stats) {
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) st_x += __shfl_xor_sync(kFull, st_x, mm);
    if (lane == 0 && st_x) atomicAdd(&J.stats[3], st_x);
  }
}

}  // namespace swb
