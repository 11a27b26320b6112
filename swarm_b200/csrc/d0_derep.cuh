// swarm_b200/csrc/d0_derep.cuh — d = 0: dereplication, the reference's `dereplicating` loop
// (/root/reference src/derep.cc:276-354) on the device.
//
// The reference walks the amplicons in index order through an open-addressing table of clusters keyed by the
// sequence's Zobrist hash and compares sequences exactly on a hash match (:297-318); the first amplicon of a class
// becomes its seed (`seqno_first`), later ones are chained behind it in index order, and size / mass / singletons
// accumulate in the bucket (:322-344).  The result does not depend on the visiting order except through "first":
//     rep(a)  = the smallest id whose (length, sequence) equals a's
//     size / mass / singletons(rep) = sums over the class
// so every amplicon can be handled by its own thread:
//   k_derep_claim   one thread per amplicon: hash the packed words (any 64-bit hash does — every match is verified),
//                   walk the table from hash & mask (the reference's bucket rule, :299); a table entry is ONE 64-bit
//                   word  tag32 | id32  (tag = high half of the hash), so a bucket step is one 8-byte load and
//                   foreign entries are skipped without touching their sequences.  Empty slot -> 64-bit CAS; same tag
//                   -> exact comparison of length and packed words, then atomicMin on the entry (same tag, smaller id
//                   wins): when the kernel ends the entry of a class holds its smallest id.  All members of a class
//                   walk the same probe sequence and entries are never removed, so a class owns exactly one slot.
//   k_derep_gather  rep[a] = id in a's slot; the class sums by atomics on the representative's counters.
// Bytes per amplicon: the packed row (8*stride) + length + abundance read once, ~1.4 table words per walk, one row
// re-read per verified match, 4 B rep written — an HBM-streaming pass plus one random 8-byte access.
#pragma once
#include "common.cuh"

namespace swb {

struct DerepParams {
  const uint64_t *words;
  const uint32_t *len;
  const uint64_t *abundance;
  uint32_t n, stride;
  unsigned long long *table;        // slots entries, ~0 = empty
  uint64_t slot_mask;
  uint32_t *slot_of;                // n: where a's class lives
  uint32_t *rep;                    // n
  unsigned long long *mass;         // n, nonzero at representatives
  uint32_t *size, *singletons;      // n
  unsigned long long *stats;        // [0] clusters [1] table steps [2] exact comparisons ([1], [2] only with count_steps)
  int count_steps;
};

__device__ __forceinline__ uint64_t derep_mix(uint64_t h, uint64_t v) {
  h ^= v;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 32;
  return h;
}

__global__ void __launch_bounds__(256) k_derep_claim(DerepParams D) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= D.n) return;
  const uint64_t *row = D.words + static_cast<uint64_t>(a) * D.stride;
  const uint32_t L = D.len[a];
  const uint32_t nw = (L + 31) >> 5;
  uint64_t h = derep_mix(0x9e3779b97f4a7c15ull, L);               // "A" and "AA" pack to the same words: the length is part of the key
  for (uint32_t k = 0; k < nw; ++k) h = derep_mix(h, row[k]);
  h ^= h >> 29; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 32;
  const unsigned long long tag = h & 0xFFFFFFFF00000000ull;
  const unsigned long long mine = tag | a;
  uint64_t idx = h & D.slot_mask;
  uint32_t steps = 0, compares = 0;
  for (;;) {
    ++steps;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&D.table[idx]);
    if (cur == ~0ull) {
      cur = atomicCAS(&D.table[idx], ~0ull, mine);
      if (cur == ~0ull) break;                                    // claimed: a is (for now) the class's representative
    }
    if ((cur & 0xFFFFFFFF00000000ull) == tag) {
      const uint32_t c = static_cast<uint32_t>(cur);
      ++compares;
      bool same = D.len[c] == L;
      if (same) {
        const uint64_t *other = D.words + static_cast<uint64_t>(c) * D.stride;
        for (uint32_t k = 0; k < nw; ++k) same = same && (other[k] == row[k]);
      }
      if (same) {
        if (a < c) atomicMin(&D.table[idx], mine);
        break;
      }
    }
    idx = (idx + 1) & D.slot_mask;
  }
  D.slot_of[a] = static_cast<uint32_t>(idx);
  if (D.count_steps) {
    atomicAdd(&D.stats[1], static_cast<unsigned long long>(steps));
    if (compares) atomicAdd(&D.stats[2], static_cast<unsigned long long>(compares));
  }
}

__global__ void __launch_bounds__(256) k_derep_gather(DerepParams D) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t r = kNone;
  unsigned long long ab = 0;
  if (a < D.n) {
    r = static_cast<uint32_t>(D.table[D.slot_of[a]]);
    D.rep[a] = r;
    ab = D.abundance[a];
  }
  // lanes of a warp that share a representative (raw reads: most of them do) add once
  const uint32_t peers = __match_any_sync(kFull, r);
  unsigned long long m = ab;
  uint32_t sz = 1, sg = ab == 1 ? 1u : 0u;
  if (__any_sync(kFull, peers != (1u << lane))) {                 // warp-uniform: every lane runs the 32 exchanges
    m = 0; sz = 0; sg = 0;
    for (int l = 0; l < 32; ++l) {
      const unsigned long long o = shfl_u64(ab, l);
      if ((peers >> l) & 1u) { m += o; ++sz; sg += o == 1 ? 1u : 0u; }
    }
  }
  if (r != kNone && lane == static_cast<uint32_t>(__ffs(peers) - 1)) {
    atomicAdd(&D.mass[r], m);
    atomicAdd(&D.size[r], sz);
    if (sg) atomicAdd(&D.singletons[r], sg);
  }
  const int heads = __syncthreads_count(r != kNone && r == a);    // one counter update per CTA
  if (threadIdx.x == 0 && heads) atomicAdd(&D.stats[0], static_cast<unsigned long long>(heads));
}

}  // namespace swb
