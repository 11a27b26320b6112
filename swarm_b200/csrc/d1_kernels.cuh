// swarm_b200/csrc/d1_kernels.cuh — d=1 hot path: index (hash + table + filter), neighbour network
// (microvariant enumeration + filter probe + bucket walk + exact verification), clustering.
//
// What each kernel replaces in the reference (/root/reference) is stated at the kernel.  Design:
//   * one WARP per seed; lane l owns the contiguous positions [l*c, (l+1)*c), c = ceil((L+1)/32);
//   * Zobrist hashing H(x) = XOR_p Z[p][x_p] (src/zobrist.cc:134-184).  The reference's rolling
//     deletion/insertion hashes (src/variants.cc:210-246) are serial along the sequence; here they
//     are three warp-level exclusive XOR scans (A_i=Z[i][s_i], B_i=Z[i-1][s_i], C_i=Z[i+1][s_i]):
//        sub(p,b) = H ^ Z[p][s_p] ^ Z[p][b]
//        del(p)   = PA(p) ^ TB ^ PB(p+1)            PA/PB/PC = exclusive prefixes, TB/TC = totals
//        ins(p,b) = PA(p) ^ Z[p][b] ^ TC ^ PC(p)
//     so every lane derives its variants' hashes in O(1) each, in registers;
//   * each variant probes one 64-bit word of an L2-resident blocked filter; survivors (~1 %) are
//     compacted with __ballot_sync into a per-warp shared-memory queue and walked 4 at a time, 8
//     lanes per survivor reading 8 consecutive 16-byte slots = one coalesced 128-byte line;
//   * hash matches are verified exactly on the packed words (funnel-shift compare);
//   * the packed seeds are staged into shared memory by per-warp double-buffered 1-D TMA bulk copies.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace swb {

constexpr int kWarpsPerCta = 8;
constexpr int kQueueCap = 64;    // filter survivors per warp
constexpr int kEdgeCap = 96;     // staged links per warp
constexpr int kMaxBatch = 8;     // seeds per TMA batch (per warp)

struct D1Params {
  const uint64_t *words;   // n_padded * stride
  const uint32_t *len;
  const uint64_t *abundance;
  uint32_t n;
  uint32_t stride;
  uint32_t batch;          // seeds per TMA batch (even)
  const uint64_t *ztab;    // zlen*4, position-major: ztab[p*4+b]
  uint32_t zlen;
  Slot *slots;
  uint64_t slot_mask;
  uint2 *filter;           // 64-bit blocks
  uint32_t filter_mask;    // blocks - 1
  uint64_t *hashes;        // n
  uint2 *edges;            // (src,dst)
  unsigned long long *edge_count;
  uint64_t edge_cap;
  unsigned long long *stats;   // [0] variants [1] filter passes [2] slots visited [3] exact compares
  uint32_t seed_begin, seed_end;
  int no_cluster_breaking;
  uint32_t *dup_flag;
  uint32_t dbg;            // experiment switches (SWB200_DEBUG env): 1 = skip bucket walks
};

__device__ __forceinline__ uint32_t base_at(const uint64_t *w, uint32_t p) {
  return static_cast<uint32_t>(w[p >> 5] >> ((p & 31u) << 1)) & 3u;
}

// =================================================================================================
// k_d1_index — "Hashing sequences" phase.  Replaces zobrist_hash (src/zobrist.cc:134-184, stored at
// src/db.cc:761), hash_insert (src/algod1.cc:188-208: first free slot by linear probing from
// (hash>>32)&mask, src/hashtable.cc:47-60) and bloom_set (src/bloompat.cc:62-65).  One warp per
// thread per amplicon; slots are claimed with a 64-bit CAS on {id,len}.
// =================================================================================================
__global__ void __launch_bounds__(256) k_d1_index(D1Params P) {
  extern __shared__ uint64_t zs[];
  for (uint32_t i = threadIdx.x; i < P.zlen * 4; i += blockDim.x) zs[i] = P.ztab[i];
  __syncthreads();
  const uint32_t nthreads = gridDim.x * blockDim.x;
  // one THREAD per amplicon: the 32 lanes of a warp walk positions in lockstep, so the Zobrist reads
  // of one step fall in one 32-byte window of shared memory (broadcast), and the 32 slot claims of a
  // warp are 32 independent atomics in flight.
  for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < P.n; a += nthreads) {
    const uint32_t L = P.len[a];
    const uint64_t *w = P.words + static_cast<uint64_t>(a) * P.stride;
    uint64_t h = 0;
    for (uint32_t j = 0; (j << 5) < L; ++j) {
      uint64_t word = w[j];
      const uint32_t p0 = j << 5;
      const uint32_t cnt = min(32u, L - p0);
      for (uint32_t q = 0; q < cnt; ++q) {
        h ^= zs[(p0 + q) * 4 + (static_cast<uint32_t>(word) & 3u)];
        word >>= 2;
      }
    }
    P.hashes[a] = h;
    uint64_t idx = (h >> 32) & P.slot_mask;
    const unsigned long long mine = (static_cast<unsigned long long>(L) << 32) | a;
    for (;;) {
      unsigned long long *cell = reinterpret_cast<unsigned long long *>(&P.slots[idx].id);
      const unsigned long long old = atomicCAS(cell, 0xFFFFFFFFFFFFFFFFull, mine);
      if (old == 0xFFFFFFFFFFFFFFFFull) { P.slots[idx].hash = h; break; }
      idx = (idx + 1) & P.slot_mask;
    }
    const uint2 pat = filter_pattern(h);
    uint32_t *blk = reinterpret_cast<uint32_t *>(P.filter + (static_cast<uint32_t>(h) & P.filter_mask));
    atomicOr(blk, pat.x);
    atomicOr(blk + 1, pat.y);
  }
}

// k_d1_dupcheck — the duplicate test of hash_insert (src/algod1.cc:174-185,193-200), done after all
// inserts are visible: an amplicon whose bucket run holds another id with the same hash, length and
// words is a duplicate (the reference then aborts, :1141-1150).  One thread per amplicon.
__global__ void __launch_bounds__(256) k_d1_dupcheck(D1Params P) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.n) return;
  const uint64_t h = P.hashes[a];
  const uint32_t L = P.len[a];
  uint64_t idx = (h >> 32) & P.slot_mask;
  for (;;) {
    const uint4 s = ld_slot(&P.slots[idx]);
    if (s.z == kNone) break;
    const uint64_t sh = (static_cast<uint64_t>(s.y) << 32) | s.x;
    if (sh == h && s.z != a && s.w == L) {
      const uint64_t *x = P.words + static_cast<uint64_t>(a) * P.stride;
      const uint64_t *y = P.words + static_cast<uint64_t>(s.z) * P.stride;
      bool same = true;
      for (uint32_t j = 0; (j << 5) < L; ++j) same = same && (x[j] == y[j]);
      if (same) atomicExch(P.dup_flag, 1u);
    }
    idx = (idx + 1) & P.slot_mask;
  }
}

// =================================================================================================
// Per-warp state of the network kernel (all in shared memory, owned by one warp).
// =================================================================================================
struct WarpScratch {
  uint64_t qhash[kQueueCap];
  uint32_t qcode[kQueueCap];
  uint2 edges[kEdgeCap];
};

template <typename SCR>
__device__ __forceinline__ void flush_edges(const D1Params &P, SCR &S, uint32_t &en, uint32_t lane) {
  if (en == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(P.edge_count, static_cast<unsigned long long>(en));
  base = shfl_u64(base, 0);
  for (uint32_t i = lane; i < en; i += 32)
    if (base + i < P.edge_cap) P.edges[base + i] = S.edges[i];
  __syncwarp();
  en = 0;
}

// word k of the variant (type,pos,base) of the seed held in `sw` (nw valid words, zero beyond) —
// the sequence generate_variant_sequence would build (src/variants.cc:78-115), one word at a time.
__device__ __forceinline__ uint64_t variant_word(const uint64_t *sw, uint32_t nw, uint32_t type, uint32_t pos,
                                                 uint32_t base, uint32_t k) {
  const uint32_t wp = pos >> 5, sh = (pos & 31u) << 1;
  const uint64_t cur = k < nw ? sw[k] : 0ull;
  if (type == 0) return k == wp ? (cur & ~(3ull << sh)) | (static_cast<uint64_t>(base) << sh) : cur;
  if (type == 1) {  // deletion of pos: everything above shifts down by one nucleotide
    if (k < wp) return cur;
    const uint64_t nxt = (k + 1 < nw) ? sw[k + 1] : 0ull;
    if (k == wp) {
      const uint64_t low = sh ? (cur & ((1ull << sh) - 1)) : 0ull;
      const uint64_t high = (sh < 62) ? ((cur >> (sh + 2)) << sh) : 0ull;
      return low | high | (nxt << 62);
    }
    return (cur >> 2) | (nxt << 62);
  }
  // insertion of `base` before pos: everything at/above pos shifts up by one nucleotide
  if (k < wp) return cur;
  if (k == wp) {
    const uint64_t low = sh ? (cur & ((1ull << sh) - 1)) : 0ull;
    const uint64_t high = (sh < 62) ? ((cur >> sh) << (sh + 2)) : 0ull;
    return low | (static_cast<uint64_t>(base) << sh) | high;
  }
  const uint64_t prv = (k - 1 < nw) ? sw[k - 1] : 0ull;
  return (cur << 2) | (prv >> 62);
}

// Walk the buckets of the queued filter survivors: find_variant_matches' bucket loop
// (src/algod1.cc:568-602) + check_variant (src/variants.cc:118-165), 4 survivors per step, 8 lanes
// each.  MODE 0 = FULL (link seed->amp under the abundance rule), 1 = HALF (pair found once; derive
// both directions).
template <int MODE, bool STATS, typename SCR>
__device__ __forceinline__ void drain_queue(const D1Params &P, SCR &S, const uint64_t *sw, uint32_t seed,
                                            uint32_t L, uint32_t &qn, uint32_t &en, uint32_t lane,
                                            unsigned long long &st_slots, unsigned long long &st_cmp) {
  const uint32_t sub = lane >> 3, j = lane & 7u;
  const uint32_t nw = (L + 31) >> 5;
  __syncwarp();
  if (P.dbg & 1u) { qn = 0; return; }
  for (uint32_t b0 = 0; b0 < qn; b0 += 4) {
    const uint32_t e = b0 + sub;
    bool done = e >= qn;
    const uint64_t h = done ? 0ull : S.qhash[e];
    const uint32_t code = done ? 0u : S.qcode[e];
    const uint32_t type = code >> 30, vbase = (code >> 28) & 3u, pos = code & 0x0FFFFFFFu;
    const uint32_t vlen = type == 0 ? L : (type == 1 ? L - 1 : L + 1);
    uint64_t idx = (h >> 32) & P.slot_mask;
    while (__any_sync(kFull, !done)) {
      if (en > kEdgeCap - 8) flush_edges(P, S, en, lane);
      uint4 s = make_uint4(0, 0, kNone, 0);
      if (!done) { s = ld_slot(&P.slots[(idx + j) & P.slot_mask]); }
      const bool empty = s.z == kNone;
      const bool match = !empty && ((static_cast<uint64_t>(s.y) << 32) | s.x) == h;
      const uint32_t be = (__ballot_sync(kFull, empty && !done) >> (sub * 8)) & 0xFFu;
      uint32_t bm = (__ballot_sync(kFull, match && !done) >> (sub * 8)) & 0xFFu;
      const uint32_t first_empty = be ? (__ffs(be) - 1) : 8u;
      bm &= (1u << first_empty) - 1u;
      if (STATS && !done && j == 0) st_slots += first_empty < 8 ? first_empty : 8;
      bool hit = false;
      // verify candidates (normally at most one per survivor); loop is warp-uniform
      while (__any_sync(kFull, bm != 0 && !hit)) {
        if (en > kEdgeCap - 8) flush_edges(P, S, en, lane);
        const bool act = bm != 0 && !hit && !done;
        const uint32_t mj = act ? (__ffs(bm) - 1) : 0u;
        const uint32_t cid = __shfl_sync(kFull, s.z, (sub << 3) + mj);
        const uint32_t clen = __shfl_sync(kFull, s.w, (sub << 3) + mj);
        bool bad = false;
        if (act) {
          if (cid == seed || clen != vlen) bad = true;
          else {
            const uint64_t *cw = P.words + static_cast<uint64_t>(cid) * P.stride;
            const uint32_t vw = (vlen + 31) >> 5;
            for (uint32_t k = j; k < vw; k += 8)
              if (variant_word(sw, nw, type, pos, vbase, k) != __ldg(cw + k)) bad = true;
          }
          if (STATS && j == 0) st_cmp += 1;
        }
        const uint32_t bb = (__ballot_sync(kFull, bad) >> (sub * 8)) & 0xFFu;
        const bool ok = act && bb == 0;
        // emit links
        bool e1 = false, e2 = false;
        uint2 l1 = make_uint2(0, 0), l2 = make_uint2(0, 0);
        if (ok && j == 0) {
          if (MODE == 0) {
            if (P.no_cluster_breaking || P.abundance[seed] >= P.abundance[cid]) { e1 = true; l1 = make_uint2(seed, cid); }
          } else {
            const uint64_t as = P.abundance[seed], ac = P.abundance[cid];
            if (P.no_cluster_breaking || as >= ac) { e1 = true; l1 = make_uint2(seed, cid); }
            if (P.no_cluster_breaking || ac >= as) { e2 = true; l2 = make_uint2(cid, seed); }
          }
        }
        const uint32_t m1 = __ballot_sync(kFull, e1);
        if (e1) S.edges[en + __popc(m1 & ((1u << lane) - 1u))] = l1;
        en += __popc(m1);
        const uint32_t m2 = __ballot_sync(kFull, e2);
        if (e2) S.edges[en + __popc(m2 & ((1u << lane) - 1u))] = l2;
        en += __popc(m2);
        if (ok) hit = true;          // first verified hit wins (src/algod1.cc:596)
        if (act) bm &= bm - 1;
      }
      if (!done) {
        if (hit || be != 0) done = true;
        else idx += 8;
      }
    }
  }
  __syncwarp();
  qn = 0;
}

// One position's microvariants (NV candidate slots): issue all filter loads, then test; survivors
// are compacted into the warp queue with __ballot_sync.
template <int MODE, int NV, bool STATS>
__device__ __forceinline__ void probe_batch(const D1Params &P, WarpScratch &S, const uint64_t *sw, uint32_t seed,
                                            uint32_t L, const bool (&vv)[NV], const uint64_t (&vh)[NV],
                                            const uint32_t (&vc)[NV], uint32_t &qn, uint32_t &en, uint32_t lane,
                                            unsigned long long &st_pass, unsigned long long &st_slots,
                                            unsigned long long &st_cmp) {
  uint2 w[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    w[k] = make_uint2(0u, 0u);
    if (vv[k]) {
      const uint2 *fp = P.filter + (static_cast<uint32_t>(vh[k]) & P.filter_mask);
      w[k] = ld_filter(fp);
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const uint2 m = filter_pattern(vh[k]);
    const bool pass = vv[k] && ((w[k].x & m.x) == m.x) && ((w[k].y & m.y) == m.y);
    const uint32_t bal = __ballot_sync(kFull, pass);
    if (bal) {
      const uint32_t cnt = __popc(bal);
      if (qn + cnt > kQueueCap) drain_queue<MODE, STATS, WarpScratch>(P, S, sw, seed, L, qn, en, lane, st_slots, st_cmp);
      if (pass) {
        const uint32_t at = qn + __popc(bal & ((1u << lane) - 1u));
        S.qhash[at] = vh[k];
        S.qcode[at] = vc[k];
      }
      qn += cnt;
      if (STATS && lane == 0) st_pass += cnt;
    }
  }
}

// Enumerate the microvariants of one seed (src/variants.cc:184-249: same set, same canonicalisation;
// the order differs, which only matters for the pre-sort order of -j rows) and hand each to F.
// F(valid, hash, code) is called warp-uniformly.
template <int MODE, typename F>
__device__ __forceinline__ void enumerate_variants(const uint64_t *zs, const uint64_t *sw, uint32_t L, uint32_t lane,
                                                   F &&emit) {
  const uint32_t c = (L + 1 + 31) >> 5;          // positions per lane, position L = "append" slot
  const uint32_t p0 = lane * c;
  const uint32_t p1 = min(p0 + c, L + 1);
  // local XORs of A_i, B_i, C_i over this lane's real positions
  uint64_t la = 0, lb = 0, lc = 0;
  for (uint32_t p = p0; p < p1 && p < L; ++p) {
    const uint32_t s = base_at(sw, p);
    la ^= zs[p * 4 + s];
    if (p >= 1) lb ^= zs[(p - 1) * 4 + s];
    if (MODE == 0) lc ^= zs[(p + 1) * 4 + s];
  }
  // warp exclusive XOR scans + totals
  uint64_t ia = la, ib = lb, ic = lc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t ta = shfl_up_u64(ia, d), tb = shfl_up_u64(ib, d);
    uint64_t tc = 0;
    if (MODE == 0) tc = shfl_up_u64(ic, d);
    if (lane >= static_cast<uint32_t>(d)) { ia ^= ta; ib ^= tb; ic ^= tc; }
  }
  const uint64_t TA = shfl_u64(ia, 31), TB = shfl_u64(ib, 31);
  uint64_t TC = 0;
  if (MODE == 0) TC = shfl_u64(ic, 31);
  uint64_t pa = ia ^ la, pb = ib ^ lb, pc = ic ^ lc;
  uint32_t prev = (p0 >= 1 && p0 - 1 < L) ? base_at(sw, p0 - 1) : 4u;

  for (uint32_t it = 0; it < c; ++it) {          // warp-uniform trip count
    const uint32_t p = p0 + it;
    const bool inr = p < p1;                     // p <= L
    const bool real = inr && p < L;
    const uint32_t s = real ? base_at(sw, p) : 4u;
    uint64_t z[4] = {0, 0, 0, 0};
    if (inr) {
#pragma unroll
      for (int b = 0; b < 4; ++b) z[b] = zs[p * 4 + b];
    }
    const uint64_t zcur = !real ? 0ull : (s == 0 ? z[0] : (s == 1 ? z[1] : (s == 2 ? z[2] : z[3])));
    const uint64_t zB = (real && p >= 1) ? zs[(p - 1) * 4 + s] : 0ull;
    constexpr int NV = MODE == 0 ? 9 : 3;
    bool vv[NV];
    uint64_t vh[NV];
    uint32_t vc[NV];
    if (MODE == 0) {
      // substitutions (3 valid of 4 slots), deletion, insertions before p
#pragma unroll
      for (uint32_t b = 0; b < 4; ++b) {
        vv[b] = real && b != s;
        vh[b] = TA ^ zcur ^ z[b];
        vc[b] = (0u << 30) | (b << 28) | p;
      }
      vv[4] = real && (p == 0 || s != prev);       // canonical deletion: first of a homopolymer run
      vh[4] = pa ^ TB ^ pb ^ zB;
      vc[4] = (1u << 30) | p;
#pragma unroll
      for (uint32_t b = 0; b < 4; ++b) {           // canonical insertion: p == 0 or base != left neighbour
        vv[5 + b] = inr && (p == 0 || b != prev);
        vh[5 + b] = pa ^ z[b] ^ TC ^ pc;
        vc[5 + b] = (2u << 30) | (b << 28) | p;
      }
    } else {
      // HALF: each unordered substitution pair {a,b} at a position is discovered from exactly one side,
      // chosen by the tournament 0->1 0->2 1->2 1->3 2->3 3->0 (out-degree 2,2,1,1: balanced lanes):
      // slot 0 probes b = s+1 (always), slot 1 probes b = s+2 (only for s < 2); slot 2 = deletion.
      {
        const uint32_t b1 = (s + 1u) & 3u, b2 = (s + 2u) & 3u;
        const uint64_t z1 = b1 == 0 ? z[0] : (b1 == 1 ? z[1] : (b1 == 2 ? z[2] : z[3]));
        const uint64_t z2 = b2 == 0 ? z[0] : (b2 == 1 ? z[1] : (b2 == 2 ? z[2] : z[3]));
        vv[0] = real;
        vh[0] = TA ^ zcur ^ z1;
        vc[0] = (0u << 30) | (b1 << 28) | p;
        vv[1] = real && s < 2u;
        vh[1] = TA ^ zcur ^ z2;
        vc[1] = (0u << 30) | (b2 << 28) | p;
      }
      vv[2] = real && (p == 0 || s != prev);
      vh[2] = pa ^ TB ^ pb ^ zB;
      vc[2] = (1u << 30) | p;
    }
    emit(vv, vh, vc);
    if (real) {
      pa ^= zcur;
      pb ^= zB;
      if (MODE == 0) pc ^= zs[(p + 1) * 4 + s];
      prev = s;
    }
  }
}

// =================================================================================================
// k_d1_network — "Building network" phase: network_thread -> check_variants -> generate_variants ->
// find_variant_matches -> check_variant (src/algod1.cc:558-670, src/variants.cc:118-249).
// Persistent warps; warp g handles seed batches g, g+G, ... of its shard; each batch of `batch`
// consecutive packed seeds is one contiguous byte range fetched by a 1-D TMA bulk copy into the
// warp's double buffer while the previous batch is being processed.
// =================================================================================================
template <int MODE, bool STATS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MODE == 0 ? 2 : 3) k_d1_network(D1Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [ztab zlen*4 u64][per-warp: 2 * batch*stride u64][per-warp WarpScratch][per-warp 2 mbarriers]
  uint64_t *zs = reinterpret_cast<uint64_t *>(smem_raw);
  const uint32_t zwords = (P.zlen * 4 + 1) & ~1u;
  const uint32_t buf_words = P.batch * P.stride;               // even number of words -> 16-byte multiple
  uint64_t *bufs = zs + zwords;
  WarpScratch *scr = reinterpret_cast<WarpScratch *>(bufs + static_cast<size_t>(kWarpsPerCta) * 2 * buf_words);
  uint64_t *bars = reinterpret_cast<uint64_t *>(scr + kWarpsPerCta);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < P.zlen * 4; i += blockDim.x) zs[i] = P.ztab[i];
  if (lane == 0) { mbar_init(&bars[warp * 2], 1); mbar_init(&bars[warp * 2 + 1], 1); }
  mbar_fence_init();
  __syncthreads();

  WarpScratch &S = scr[warp];
  uint64_t *mybuf = bufs + static_cast<size_t>(warp) * 2 * buf_words;
  uint64_t *mybar = &bars[warp * 2];
  const uint32_t G = gridDim.x * kWarpsPerCta;
  const uint32_t g = blockIdx.x * kWarpsPerCta + warp;
  const uint32_t first_batch = P.seed_begin / P.batch;                       // seed_begin is batch aligned
  const uint32_t n_batches = (P.seed_end - P.seed_begin + P.batch - 1) / P.batch;
  const uint32_t buf_bytes = buf_words * 8;

  unsigned long long st_var = 0, st_pass = 0, st_slots = 0, st_cmp = 0;
  uint32_t qn = 0, en = 0;
  uint32_t phase0 = 0, phase1 = 0;

  uint32_t bi = g;
  if (bi < n_batches && lane == 0) {
    mbar_expect_tx(&mybar[0], buf_bytes);
    tma_load_1d(mybuf, P.words + static_cast<uint64_t>(first_batch + bi) * buf_words, buf_bytes, &mybar[0]);
  }
  uint32_t cur = 0;
  for (; bi < n_batches; bi += G) {
    const uint32_t nb = bi + G;
    if (nb < n_batches && lane == 0) {           // prefetch the next batch into the other buffer
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
      const uint64_t *sw = tile + k * P.stride;
      constexpr int NV = MODE == 0 ? 9 : 3;
      auto emit = [&](const bool (&vv)[NV], const uint64_t (&vh)[NV], const uint32_t (&vc)[NV]) {
        if (STATS) {
#pragma unroll
          for (int k = 0; k < NV; ++k) st_var += vv[k] ? 1u : 0u;
        }
        probe_batch<MODE, NV, STATS>(P, S, sw, seed, L, vv, vh, vc, qn, en, lane, st_pass, st_slots, st_cmp);
      };
      enumerate_variants<MODE>(zs, sw, L, lane, emit);
      if (qn) drain_queue<MODE, STATS, WarpScratch>(P, S, sw, seed, L, qn, en, lane, st_slots, st_cmp);
    }
    __syncwarp();                                // all lanes done reading the tile before it is re-filled
    cur ^= 1;
  }
  flush_edges(P, S, en, lane);
  if (STATS && P.stats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      st_var += __shfl_xor_sync(kFull, st_var, m);
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

// debug: dump the variants of one seed (test hook, swb200_debug_variants)
template <int MODE>
__global__ void k_d1_debug_variants(D1Params P, uint32_t seed, uint64_t *out_hash, uint32_t *out_code, uint32_t cap,
                                    uint32_t *count) {
  extern __shared__ uint64_t zs[];
  for (uint32_t i = threadIdx.x; i < P.zlen * 4; i += blockDim.x) zs[i] = P.ztab[i];
  __shared__ uint64_t sw[1024];
  for (uint32_t i = threadIdx.x; i < P.stride && i < 1024; i += blockDim.x)
    sw[i] = P.words[static_cast<uint64_t>(seed) * P.stride + i];
  __syncthreads();
  const uint32_t lane = threadIdx.x;
  uint32_t qn = 0;
  constexpr int NV = MODE == 0 ? 9 : 3;
  auto emit = [&](const bool (&vv)[NV], const uint64_t (&vh)[NV], const uint32_t (&vc)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const uint32_t bal = __ballot_sync(kFull, vv[k]);
      if (vv[k]) {
        const uint32_t at = qn + __popc(bal & ((1u << lane) - 1u));
        if (at < cap) { out_hash[at] = vh[k]; out_code[at] = vc[k]; }
      }
      qn += __popc(bal);
    }
  };
  enumerate_variants<MODE>(zs, sw, P.len[seed], lane, emit);
  if (lane == 0) *count = qn;
}

// =================================================================================================
// Clustering.  The reference's greedy loop (src/algod1.cc:1185-1280) is sequential over seeds; its
// result has a closed form (checked against the oracle in tests):
//   swarm(v)  = the smallest amplicon id that reaches v through directed links;
//   generation(v), parent(v) = BFS depth from the swarm's seed over links whose two ends are in the
//   same swarm, and the smallest-id in-swarm predecessor one level up (process_seed :673-718: the
//   members of a generation are scanned in id order and the first claimer wins).
// k_label_* iterate min-label propagation (+ pointer jumping: label[label[v]] also reaches v);
// k_bfs_relax iterates a 64-bit atomicMin on key = generation<<32 | parent.
// =================================================================================================
__global__ void k_label_init(uint32_t *label, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) label[i] = i;
}
__global__ void k_label_edges(const uint2 *edges, uint64_t m, uint32_t *label, uint32_t *changed) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  bool ch = false;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < m; e += stride) {
    const uint2 ed = edges[e];
    const uint32_t ls = label[ed.x];
    if (ls < label[ed.y]) { atomicMin(&label[ed.y], ls); ch = true; }
  }
  if (__any_sync(kFull, ch) && (threadIdx.x & 31) == 0) *changed = 1;
}
__global__ void k_label_jump(uint32_t *label, uint32_t n, uint32_t *changed) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  bool ch = false;
  if (v < n) {
    const uint32_t l = label[v];
    const uint32_t ll = label[l];
    if (ll < l) { atomicMin(&label[v], ll); ch = true; }
  }
  if (__any_sync(kFull, ch) && (threadIdx.x & 31) == 0) *changed = 1;
}
__global__ void k_bfs_init(const uint32_t *label, unsigned long long *key, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (label[i] == i) ? 0x00000000FFFFFFFFull : 0xFFFFFFFFFFFFFFFFull;
}
__global__ void k_bfs_relax(const uint2 *edges, uint64_t m, const uint32_t *label, unsigned long long *key,
                            uint32_t *changed) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  bool ch = false;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < m; e += stride) {
    const uint2 ed = edges[e];
    if (label[ed.x] != label[ed.y]) continue;
    const unsigned long long ku = key[ed.x];
    if (ku == 0xFFFFFFFFFFFFFFFFull) continue;
    const unsigned long long cand = (((ku >> 32) + 1) << 32) | ed.x;
    if (cand < key[ed.y]) { atomicMin(&key[ed.y], cand); ch = true; }
  }
  if (__any_sync(kFull, ch) && (threadIdx.x & 31) == 0) *changed = 1;
}
// ---- fused formulation (default): ONE relaxation on key = label<<32 | generation.
// key[v] = min over links u->v of key[u]+1 (lexicographic: the smallest reaching id wins; among
// predecessors carrying that id, the smallest depth), iterated to the fixed point with 64-bit atomicMin;
// then parent[v] = min { u : u->v, key[u]+1 == key[v] }.  Same result as k_label_* + k_bfs_* in half the
// passes over the link list.
__global__ void k_key_init(unsigned long long *key, uint32_t *parent, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = static_cast<unsigned long long>(i) << 32; parent[i] = kNone; }
}
__global__ void k_key_relax(const uint2 *edges, uint64_t m, unsigned long long *key, uint32_t *changed) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  bool ch = false;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < m; e += stride) {
    const uint2 ed = edges[e];
    const unsigned long long cand = key[ed.x] + 1ull;
    if (cand < key[ed.y]) { atomicMin(&key[ed.y], cand); ch = true; }
  }
  if (__any_sync(kFull, ch) && (threadIdx.x & 31) == 0) *changed = 1;
}
__global__ void k_key_parent(const uint2 *edges, uint64_t m, const unsigned long long *key, uint32_t *parent) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < m; e += stride) {
    const uint2 ed = edges[e];
    if (key[ed.x] + 1ull == key[ed.y]) atomicMin(&parent[ed.y], ed.x);
  }
}
__global__ void k_key_unpack(const unsigned long long *key, uint32_t *label, uint32_t *generation, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { label[i] = static_cast<uint32_t>(key[i] >> 32); generation[i] = static_cast<uint32_t>(key[i]); }
}

// ---- the same fused relaxation as ONE persistent cooperative kernel (default): init, every round, the parent
// pass and the unpacking run back to back with grid-wide barriers in between, so a clustering costs one launch
// and no host round trip (the host-looped variant above pays a launch + a D2H flag read per round, ~15 rounds).
// flags[3] rotate: round r raises flags[r%3]; flags[(r+1)%3] is cleared during round r (its last readers
// passed the barrier of round r-1).
// A link u->v can only lower key[v] in round r if key[u] was lowered in round r-1, so every round tests one bit
// of a "lowered last round" bitmap (n/8 bytes, L2-resident) before touching the two random 8-byte keys of the
// link: after the first two rounds most links are skipped (the first cut relaxed every link in every round and
// was bound by ~300 M random L2 sectors: 1.2 ms at 10 M amplicons).  bits = 3 rotating bitmaps of `nwords` words:
// round r reads bits[r%3], sets bits[(r+1)%3], clears bits[(r+2)%3].
// HINT: the link list (70 MB at 10 M amplicons, re-read every round) is loaded with the streaming policy (ld.global.cs,
// evict-first) and the per-amplicon outputs are stored with st.global.cs, so that they do not push the randomly accessed
// key[] (80 MB) out of the 126 MB L2: the working set of this kernel sits right at the L2's capacity, and without the
// hints the same binary measured 1.1 ms on one box and 2.7 ms on another (gpurun_out/r1s vs r1u).
constexpr int kClU = 8;
// PACK (default whenever 2 * id_bits + 10 <= 64, i.e. up to 2^27 amplicons): the relaxed word is
//   swarm << (gb + ib) | generation << ib | parent        (ib = bits of an amplicon id, gb = min(32, 64 - 2 ib))
// and the offer of link u -> v is ((word[u] >> ib) + 1) << ib | u, so ONE 64-bit atomicMin settles the parent together with the
// key: the lexicographic minimum over (swarm, generation, u) is exactly parent[v] = min { u : u -> v, key[u] + 1 == key[v] }
// (every in-neighbour's LAST offer carries its final key, earlier ones were larger).  The separate parent pass — the link list and
// two random keys per link once more, 0.13 of 1.1 ms at 10 M — disappears.  A vertex is marked "lowered" only when its KEY part
// went down (a smaller parent under the same key changes nothing for its out-links).  A swarm deeper than gb bits hold (a final
// generation of 2^gb - 1 or more) raises bit 31 of *rounds_out and the host runs the unpacked kernel instead (engine.cu: run_cluster).
// (Tried and dropped, r2x: a coarse "any of these 128 amplicons lowered?" bitmap copied into shared memory in front of the fine
// one — fewer random bitmap reads in the late rounds, but the extra grid barrier and fold per round cost more: 1.17 vs 1.04 ms.)
template <bool HINT, bool PACK>
__global__ void __launch_bounds__(256, 4) k_cluster_persistent(const uint2 *edges, uint64_t m, unsigned long long *key, uint32_t *parent,
                                                            uint32_t *label, uint32_t *generation, uint32_t n,
                                                            volatile uint32_t *flags, uint32_t *rounds_out, uint32_t *bits,
                                                            uint32_t nwords, uint32_t ib, uint32_t gb) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned long long idmask = PACK ? (1ull << ib) - 1ull : 0ull;
  const unsigned long long gmask = PACK ? (gb >= 32 ? 0xFFFFFFFFull : (1ull << gb) - 1ull) : 0ull;
  for (uint64_t v = tid; v < n; v += nth) {
    if (PACK) key[v] = (static_cast<unsigned long long>(v) << (gb + ib)) | idmask;
    else {
      key[v] = static_cast<unsigned long long>(v) << 32;
      if (HINT) __stcs(&parent[v], kNone); else parent[v] = kNone;
    }
  }
  for (uint64_t w = tid; w < 3ull * nwords; w += nth) bits[w] = 0;
  if (tid == 0) { flags[0] = 0; flags[1] = 0; flags[2] = 0; flags[3] = 0; }
  grid.sync();
  uint32_t round = 0;
  for (;; ++round) {
    const uint32_t *rd = bits + static_cast<size_t>(round % 3) * nwords;
    uint32_t *wr = bits + static_cast<size_t>((round + 1) % 3) * nwords;
    uint32_t *cl = bits + static_cast<size_t>((round + 2) % 3) * nwords;
    if (tid == 0) flags[(round + 1) % 3] = 0;
    if (round) for (uint64_t w = tid; w < nwords; w += nth) cl[w] = 0;
    int ch = 0;
    // kClU links per thread and step, loads issued stage by stage (links, bitmap words, keys): the loop is a chain
    // of dependent random reads, and one link at a time left each warp with a single request in flight
    for (uint64_t base = tid; base < m; base += nth * kClU) {
      uint2 ed[kClU];
      bool act[kClU];
#pragma unroll
      for (int k = 0; k < kClU; ++k) {
        const uint64_t e = base + static_cast<uint64_t>(k) * nth;
        act[k] = e < m;
        ed[k] = act[k] ? (HINT ? __ldcs(&edges[e]) : edges[e]) : make_uint2(0u, 0u);
      }
      if (round) {
        uint32_t w[kClU];
#pragma unroll
        for (int k = 0; k < kClU; ++k) w[k] = act[k] ? rd[ed[k].x >> 5] : 0u;
#pragma unroll
        for (int k = 0; k < kClU; ++k) act[k] = act[k] && ((w[k] >> (ed[k].x & 31u)) & 1u);
      }
      unsigned long long ks[kClU], kd[kClU];
#pragma unroll
      for (int k = 0; k < kClU; ++k) {
        // round 0: every source still has its initial word — no load
        if (round == 0) ks[k] = PACK ? ((static_cast<unsigned long long>(ed[k].x) << (gb + ib)) | idmask) : (static_cast<unsigned long long>(ed[k].x) << 32);
        else ks[k] = act[k] ? key[ed[k].x] : ~0ull;
        kd[k] = act[k] ? key[ed[k].y] : 0ull;
      }
#pragma unroll
      for (int k = 0; k < kClU; ++k) {
        if (PACK) {
          // (word | idmask) + 1 = the key part plus one with an empty parent field; a generation that wraps into the swarm field
          // is caught when the words are unpacked (gen == gmask: nothing reachable from it can beat its legitimate offers)
          const unsigned long long cand = ((ks[k] | idmask) + 1ull) | ed[k].x;
          if (act[k] && cand < kd[k]) {
            // fire and forget (RED.MIN): waiting for the atomic's return value was 30 % of the kernel's stall samples
            // (profiles/r2y_k_cluster_persistent.txt).  "Lowered" is decided from the word loaded before: whenever this offer or a
            // concurrent better one lowers the key, kd's key was above the offer's — a superset of the true set whose extra
            // members were lowered (and marked) by somebody else in this same round.
            atomicMin(&key[ed[k].y], cand);
            if ((kd[k] | idmask) > (cand | idmask)) { atomicOr(&wr[ed[k].y >> 5], 1u << (ed[k].y & 31u)); ch = 1; }
          }
        } else {
          const unsigned long long cand = ks[k] + 1ull;
          if (act[k] && cand < kd[k]) {
            atomicMin(&key[ed[k].y], cand);
            atomicOr(&wr[ed[k].y >> 5], 1u << (ed[k].y & 31u));
            ch = 1;
          }
        }
      }
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) flags[round % 3] = 1;
    grid.sync();
    if (flags[round % 3] == 0) break;
  }
  if (!PACK) {
    for (uint64_t base = tid; base < m; base += nth * kClU) {
      uint2 ed[kClU];
      unsigned long long ks[kClU], kd[kClU];
#pragma unroll
      for (int k = 0; k < kClU; ++k) {
        const uint64_t e = base + static_cast<uint64_t>(k) * nth;
        ed[k] = e < m ? (HINT ? __ldcs(&edges[e]) : edges[e]) : make_uint2(kNone, kNone);
      }
#pragma unroll
      for (int k = 0; k < kClU; ++k) {
        ks[k] = ed[k].x != kNone ? key[ed[k].x] : 0ull;
        kd[k] = ed[k].x != kNone ? key[ed[k].y] : 0ull;
      }
#pragma unroll
      for (int k = 0; k < kClU; ++k)
        if (ed[k].x != kNone && ks[k] + 1ull == kd[k]) atomicMin(&parent[ed[k].y], ed[k].x);
    }
  }
  for (uint64_t v = tid; v < n; v += nth) {
    const unsigned long long kv = key[v];
    uint32_t lab, gen;
    if (PACK) {
      lab = static_cast<uint32_t>(kv >> (gb + ib));
      gen = static_cast<uint32_t>((kv >> ib) & gmask);
      if (gen == gmask) flags[3] = 1;                     // deeper than the packed word holds
      const uint32_t par = (kv & idmask) == idmask ? kNone : static_cast<uint32_t>(kv & idmask);
      if (HINT) __stcs(&parent[v], par); else parent[v] = par;
    } else {
      lab = static_cast<uint32_t>(kv >> 32);
      gen = static_cast<uint32_t>(kv);
    }
    if (HINT) { __stcs(&label[v], lab); __stcs(&generation[v], gen); }
    else { label[v] = lab; generation[v] = gen; }
  }
  if (PACK) {
    grid.sync();
    if (tid == 0 && rounds_out) *rounds_out = (round + 1) | (flags[3] ? 0x80000000u : 0u);
  } else if (tid == 0 && rounds_out) *rounds_out = round + 1;
}

__global__ void k_bfs_unpack(const unsigned long long *key, uint32_t *generation, uint32_t *parent, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    generation[i] = static_cast<uint32_t>(key[i] >> 32);
    parent[i] = static_cast<uint32_t>(key[i]);
  }
}

}  // namespace swb
