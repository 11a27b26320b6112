// swarm_b200/csrc/d1_tilejoin.cuh — the d=1 neighbour network as a radix-partitioned pigeonhole join that
// runs in shared memory (enum_mode JOIN, default kernel since r1k).
//
// Same mathematics as d1_join.cuh: v is a microvariant of u  <=>  ed(u, v) = 1 (src/variants.cc:184-249
// enumerates exactly the sequences one edit away), and one edit cannot touch both the first K and the last
// K nucleotides, so every linked pair shares its K-mer prefix (anchored at the start) or its K-mer suffix
// (anchored at the end).  d1_join.cuh kept the two K-mer entries per amplicon in a global open-addressing
// multimap and paid, per amplicon, two dependent chains of random 32-byte bucket reads plus two random
// packed-sequence reads per candidate pair (36 M pairs at 10 M amplicons: the rows of a dense group were
// fetched once per PAIR).  Here the join is partitioned first, the way a GPU hash join is:
//   k_tile_partition<COUNT>   every amplicon hashes its two pieces; the high bits of a piece hash pick one
//                             of T tiles (~cmax/2 entries each); pass 1 counts, k_tile_scan turns the counts
//                             into exact offsets, pass 2 appends `tag | piece | abh7 | len13 | id` entries.
//                             All entries with the same key land in the same tile.
//   k_tile_join               one CTA per tile: entries -> shared memory, counting-sorted by key, every entry
//                             that has a same-key partner gathers its packed sequence ONCE into shared memory,
//                             and all same-key pairs are then decided exactly from shared memory, one pair per
//                             thread (edit_class = the lane-parallel check_variant, src/variants.cc:118-165).
//                             HBM sees each entry once and each row at most twice.
//   k_tile_join_big           tiles that do not fit (a huge group sharing one K-mer) are swept pairwise from
//                             global memory — exact, just slower.
// Links are derived from the two abundances exactly as before (src/algod1.cc:580-583); ed = 0 pairs are the
// reference's duplicate fatal (:1141-1150).  Multi-GPU: a rank owns the tile range [t_lo, t_hi) — the hash
// range sharding of SURVEY.md §8e — scans all amplicons, keeps only its tiles' entries.
#pragma once
#include "d1_join.cuh"

namespace swb {

constexpr uint32_t kTjOutCap = 512;      // links staged per tile before one global atomicAdd
constexpr uint32_t kTjBuckets = 1024;    // counting-sort buckets per tile
constexpr uint32_t kTjMaxC = 768;        // most entries a tile can hold in shared memory (3 per thread)

struct TileJoinParams {
  const uint64_t *words;
  const uint32_t *len;
  const uint64_t *abundance;
  uint32_t n, stride, K;
  uint32_t n_tiles;                 // global tile count (the hash -> tile map is the same on every rank)
  uint32_t t_lo, t_hi;              // this rank's tiles
  uint32_t cmax;                    // entries a tile may hold to be joined in shared memory (<= kTjMaxC)
  uint32_t id_bits;                 // entry layout: key | abh(7) | len(13) | id(id_bits)
  int sorted_desc;                  // the database is sorted by abundance, descending (the reference's order, src/db.cc:392-406)
  uint32_t *tile_count;             // [t_hi - t_lo]
  unsigned long long *tile_off;     // [t_hi - t_lo + 1]
  uint32_t *tile_cursor;            // [t_hi - t_lo]
  uint32_t *big_tiles;              // local ids of the tiles with more than cmax entries
  uint32_t *big_count;
  unsigned long long *entries;
  uint2 *edges;
  unsigned long long *edge_count;
  uint64_t edge_cap;
  int ncb;
  uint32_t *dup_flag;
  uint2 *plist_ent;                 // multi-GPU: (entry lo, entry hi) and ...
  uint32_t *plist_tile;             // ... local tile of every piece this rank owns, appended by the counting pass
  unsigned long long *plist_n;      // kPlistSubs counters, kPlistPad words apart: CTA b appends to sub-list b % kPlistSubs
  uint64_t plist_cap;               // capacity of ONE sub-list
  unsigned long long *stats;        // [0] entries joined [1] same-key pairs [2] pairs enumerated [3] exact comparisons [4] rows gathered
};

// entry = key | abh(7) | len(13) | id(id_bits), key = tag | piece in the remaining top bits (>= 12).  abh is a
// 7-bit hash of the abundance: in a database sorted by abundance the smaller id always links to the larger one
// and the reverse link exists only for equal abundances — different abh proves inequality without touching
// the abundance array.
__device__ __forceinline__ unsigned long long tj_pack(const TileJoinParams &J, uint64_t h, uint32_t piece, uint32_t L, uint64_t ab,
                                                      uint32_t id) {
  const uint32_t abh = static_cast<uint32_t>((ab * 0x9E3779B97F4A7C15ull) >> 57);
  const uint32_t low_bits = J.id_bits + 20;
  const uint64_t key = (h << 1) | piece;               // truncated by the shift below
  return (key << low_bits) | (static_cast<unsigned long long>(abh) << (J.id_bits + 13)) |
         (static_cast<unsigned long long>(L & 0x1FFFu) << J.id_bits) | id;
}
__device__ __forceinline__ uint32_t tj_key(const TileJoinParams &J, unsigned long long e) { return static_cast<uint32_t>(e >> (J.id_bits + 20)); }
__device__ __forceinline__ uint32_t tj_abh(const TileJoinParams &J, unsigned long long e) { return static_cast<uint32_t>(e >> (J.id_bits + 13)) & 0x7Fu; }
__device__ __forceinline__ uint32_t tj_len(const TileJoinParams &J, unsigned long long e) { return static_cast<uint32_t>(e >> J.id_bits) & 0x1FFFu; }
__device__ __forceinline__ uint32_t tj_id(const TileJoinParams &J, unsigned long long e) {
  return static_cast<uint32_t>(e & ((1ull << J.id_bits) - 1ull));
}
__device__ __forceinline__ bool tj_compatible(const TileJoinParams &J, unsigned long long e, unsigned long long f) {
  const uint32_t Le = tj_len(J, e), Lf = tj_len(J, f);
  return tj_key(J, e) == tj_key(J, f) && Le + 1 >= Lf && Lf + 1 >= Le;
}

// pass 1 (COUNT) and pass 2 (!COUNT) of the partitioning: one thread per amplicon, both pieces
template <bool COUNT>
__global__ void __launch_bounds__(256) k_tile_partition(TileJoinParams J) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= J.n) return;
  const uint64_t *w = J.words + static_cast<uint64_t>(a) * J.stride;
  const uint32_t L = J.len[a];
  const uint64_t ab = COUNT ? 0ull : J.abundance[a];
#pragma unroll
  for (uint32_t piece = 0; piece < 2; ++piece) {
    const uint64_t h = piece_hash(w, J.stride, piece ? L - J.K : 0u, J.K, piece);
    const uint32_t tile = static_cast<uint32_t>(__umul64hi(h, static_cast<uint64_t>(J.n_tiles)));
    if (tile < J.t_lo || tile >= J.t_hi) continue;           // another rank owns this hash range
    const uint32_t t = tile - J.t_lo;
    if (COUNT) {
      atomicAdd(&J.tile_count[t], 1u);
    } else {
      const uint32_t pos = atomicAdd(&J.tile_cursor[t], 1u);
      J.entries[J.tile_off[t] + pos] = tj_pack(J, h, piece, L, ab, a);
    }
  }
}

// the kept pieces go to kPlistSubs sub-lists with their own counters: one counter for every warp of the scan serialised
// on a single L2 address (index at 2 x 10 M: 0.67 -> 1.10 ms, gpurun_out/r1n_call7)
constexpr uint32_t kPlistSubs = 64, kPlistPad = 16;

// Multi-GPU flavour of the two passes: a rank owns 1/world of the tiles but has to hash ALL amplicons to find its
// pieces, so the counting pass also appends the pieces it keeps to a list and the second pass scatters that list
// instead of hashing everything again (index at 8 x 10 M: the part that does not scale).
__global__ void __launch_bounds__(256) k_tile_partition_list(TileJoinParams J) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long e[2] = {0, 0};
  uint32_t t[2] = {kNone, kNone};
  if (a < J.n) {
    const uint64_t *w = J.words + static_cast<uint64_t>(a) * J.stride;
    const uint32_t L = J.len[a];
    const uint64_t ab = J.abundance[a];
#pragma unroll
    for (uint32_t piece = 0; piece < 2; ++piece) {
      const uint64_t h = piece_hash(w, J.stride, piece ? L - J.K : 0u, J.K, piece);
      const uint32_t tile = static_cast<uint32_t>(__umul64hi(h, static_cast<uint64_t>(J.n_tiles)));
      if (tile < J.t_lo || tile >= J.t_hi) continue;
      t[piece] = tile - J.t_lo;
      e[piece] = tj_pack(J, h, piece, L, ab, a);
      atomicAdd(&J.tile_count[t[piece]], 1u);
    }
  }
  const uint32_t b0 = __ballot_sync(kFull, t[0] != kNone), b1 = __ballot_sync(kFull, t[1] != kNone);
  const uint32_t tot = __popc(b0) + __popc(b1);
  if (tot == 0) return;
  unsigned long long base = 0;
  const uint32_t sub = blockIdx.x % kPlistSubs;
  if (lane == 0) base = atomicAdd(&J.plist_n[sub * kPlistPad], static_cast<unsigned long long>(tot));
  base = shfl_u64(base, 0);
  const uint32_t lt = (1u << lane) - 1u;
  const unsigned long long i0 = base + __popc(b0 & lt), i1 = base + __popc(b0) + __popc(b1 & lt);
  const uint64_t at = static_cast<uint64_t>(sub) * J.plist_cap;
  if (t[0] != kNone && i0 < J.plist_cap) { J.plist_ent[at + i0] = make_uint2(static_cast<uint32_t>(e[0]), static_cast<uint32_t>(e[0] >> 32)); J.plist_tile[at + i0] = t[0]; }
  if (t[1] != kNone && i1 < J.plist_cap) { J.plist_ent[at + i1] = make_uint2(static_cast<uint32_t>(e[1]), static_cast<uint32_t>(e[1] >> 32)); J.plist_tile[at + i1] = t[1]; }
}

// grid (ceil(longest sub-list / 256), kPlistSubs)
__global__ void __launch_bounds__(256) k_tile_scatter_list(TileJoinParams J) {
  const uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= J.plist_n[blockIdx.y * kPlistPad]) return;
  const uint64_t i = static_cast<uint64_t>(blockIdx.y) * J.plist_cap + k;
  const uint32_t t = J.plist_tile[i];
  const uint2 e = J.plist_ent[i];
  const uint32_t pos = atomicAdd(&J.tile_cursor[t], 1u);
  J.entries[J.tile_off[t] + pos] = (static_cast<unsigned long long>(e.y) << 32) | e.x;
}

// exclusive scan of the tile counts (one CTA) + the list of oversize tiles
__global__ void __launch_bounds__(1024) k_tile_scan(TileJoinParams J) {
  __shared__ unsigned long long part[1024];
  const uint32_t T = J.t_hi - J.t_lo;
  const uint32_t per = (T + 1023u) / 1024u;
  const uint32_t b = min(T, threadIdx.x * per), e = min(T, b + per);
  unsigned long long s = 0;
  for (uint32_t i = b; i < e; ++i) s += J.tile_count[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    const unsigned long long v = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = part[threadIdx.x] - s;
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t c = J.tile_count[i];
    J.tile_off[i] = run;
    if (c > J.cmax) J.big_tiles[atomicAdd(J.big_count, 1u)] = i;
    run += c;
  }
  if (threadIdx.x == 1023) J.tile_off[T] = part[1023];
}

// Exact test of one pair on packed words, WITHOUT data-dependent branches: 0 = identical, 1 = exactly one edit apart,
// 2 = further.  Lengths differ by at most one.  x = the longer sequence, y the other, both zero padded:
//   a_k = x_k ^ y_k                      -> Hamming distance (equal lengths) and P = first differing position
//   b_k = (x shifted down one nt)_k ^ y_k -> Q = last position where x with one nucleotide removed EARLIER still differs
// Equal lengths: one edit <=> Hamming distance 1.  Lengths Lx = Ly+1: deleting x[p] gives y <=> x[0,p) = y[0,p) and
// x[p+1+i] = y[p+i] for all i, i.e. some p <= P has every dirty b position below it <=> Q < P (Q = -1: b is clean).
// This is the lane-parallel check_variant (src/variants.cc:118-165); one uniform loop for substitutions and indels —
// the earlier two-path version (early-exit Hamming / first-mismatch-then-shift) ran with 5-17 active lanes per warp.
__device__ __forceinline__ int tj_classify(const uint64_t *ri, uint32_t Li, const uint64_t *rj, uint32_t Lj, uint32_t stride,
                                           uint64_t kmask0, uint64_t kmask1, bool &pfx_eq) {
  const bool swap = Li < Lj;
  const uint64_t *x = swap ? rj : ri, *y = swap ? ri : rj;
  // the loop only remembers the FIRST dirty word of a and the LAST dirty word of b (selects, no bit scans: the 64-bit
  // ffs / clz per word of the first version cost ~70 instructions a word); the two bit positions are found once, after it
  uint32_t ham = 0, ja = 0xFFFFFFFFu, jb = 0;
  uint64_t aw = 0, bw = 0, xv = x[0], d0 = 0, d1 = 0;
  for (uint32_t k = 0; k < stride; ++k) {
    const uint64_t xn = (k + 1 < stride) ? x[k + 1] : 0ull;
    const uint64_t yv = y[k];
    const uint64_t a = xv ^ yv;
    if (k == 0) d0 = a;
    if (k == 1) d1 = a;
    ham += __popcll((a | (a >> 1)) & 0x5555555555555555ull);
    if (a != 0 && ja == 0xFFFFFFFFu) { ja = k; aw = a; }
    const uint64_t b = ((xv >> 2) | (xn << 62)) ^ yv;
    if (b != 0) { jb = k; bw = b; }
    xv = xn;
  }
  pfx_eq = (d0 & kmask0) == 0 && (d1 & kmask1) == 0;
  if (Li == Lj) return ham > 1 ? 2 : static_cast<int>(ham);
  if (bw == 0 || aw == 0) return 1;                       // b clean: drop x[0]; a clean: x = y + one trailing base
  const uint32_t P = (ja << 5) + (static_cast<uint32_t>(__ffsll(static_cast<long long>((aw | (aw >> 1)) & 0x5555555555555555ull)) - 1) >> 1);
  const uint32_t Q = (jb << 5) + (static_cast<uint32_t>(63 - __clzll(static_cast<long long>((bw | (bw >> 1)) & 0x5555555555555555ull))) >> 1);
  return Q < P ? 1 : 2;
}

// decide one same-key pair and build its links (0, 1 or 2) in registers; rows are pointers to packed words
// (shared or global memory)
template <bool STATS>
__device__ __forceinline__ uint32_t tj_pair(const TileJoinParams &J, unsigned long long ei, unsigned long long ej, const uint64_t *ri,
                                            const uint64_t *rj, uint64_t kmask0, uint64_t kmask1, unsigned long long &st_x, uint2 &l0,
                                            uint2 &l1) {
  bool pfx_eq;
  const int cls = tj_classify(ri, tj_len(J, ei), rj, tj_len(J, ej), J.stride, kmask0, kmask1, pfx_eq);
  // a pair that shares the prefix belongs to its prefix tile; a prefix-tile pair whose first K nucleotides differ (the truncated
  // keys collided) belongs to its suffix tile: every linked pair is emitted exactly once
  if (((tj_key(J, ei) & 1u) != 0) == pfx_eq) return 0;
  if (STATS) st_x++;
  if (cls == 0) atomicExch(J.dup_flag, 1u);
  if (cls != 1) return 0;
  uint32_t a = tj_id(J, ei), v = tj_id(J, ej);
  if (J.ncb) { l0 = make_uint2(a, v); l1 = make_uint2(v, a); return 2; }
  if (J.sorted_desc) {                                   // ids ascend as abundances descend: min(a,v) -> max(a,v) always exists
    if (a > v) { const uint32_t t_ = a; a = v; v = t_; }
    l0 = make_uint2(a, v);
    if (tj_abh(J, ei) == tj_abh(J, ej) && J.abundance[a] == J.abundance[v]) { l1 = make_uint2(v, a); return 2; }
    return 1;
  }
  const uint64_t aa = J.abundance[a], av = J.abundance[v];
  uint32_t nl = 0;
  if (aa >= av) { l0 = make_uint2(a, v); nl = 1; }
  if (av >= aa) { if (nl) l1 = make_uint2(v, a); else l0 = make_uint2(v, a); ++nl; }
  return nl;
}

// exclusive scan of vals[0..N) in shared memory by 256 threads, PER consecutive items per thread (N <= 256*PER);
// leaves the total in vals[N].  Ends with a barrier.
template <int PER>
__device__ __forceinline__ void tj_block_scan(uint32_t *vals, uint32_t N, uint32_t *warp_tot, uint32_t tid) {
  uint32_t v[PER], sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t idx = tid * PER + k;
    v[k] = idx < N ? vals[idx] : 0u;
    sum += v[k];
  }
  const uint32_t lane = tid & 31u, w = tid >> 5;
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t_ = __shfl_up_sync(kFull, inc, d);
    if (lane >= static_cast<uint32_t>(d)) inc += t_;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (uint32_t j = 0; j < 8; ++j) base += (j < w) ? warp_tot[j] : 0u;
  uint32_t run = base + inc - sum;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t idx = tid * PER + k;
    if (idx < N) vals[idx] = run;
    run += v[k];
  }
  if (tid == 255) vals[N] = base + inc;
  __syncthreads();
}

// One CTA per tile.  (1) counting sort of the tile's entries by the low 10 bits of the key (shared-memory
// atomics give the rank, a block scan the offsets): same-key entries become contiguous.  (2) every entry that
// has a same-key, length-compatible partner brings its packed sequence into shared memory, once, and
// announces how many entries follow it in its bucket run.  (3) a block scan of those counts numbers all
// pairs of the tile; thread p finds pair p by binary search, so the exact comparisons run converged, one pair
// per thread, whatever the group sizes (the first cuts of this kernel walked chains per entry and decided
// pairs inside the walk: 2-6 active lanes per warp, 4.7 and 2.6 ms instead of < 1).
template <bool STATS>
__global__ void __launch_bounds__(256) k_tile_join(TileJoinParams J) {
  extern __shared__ __align__(16) unsigned char tj_smem[];
  unsigned long long *ent = reinterpret_cast<unsigned long long *>(tj_smem);
  uint64_t *rows = reinterpret_cast<uint64_t *>(ent + J.cmax);
  uint2 *out = reinterpret_cast<uint2 *>(rows + static_cast<size_t>(J.cmax) * J.stride);
  uint32_t *boff = reinterpret_cast<uint32_t *>(out + kTjOutCap);      // kTjBuckets + 1 (+1 pad)
  uint32_t *pref = boff + kTjBuckets + 2;                               // cmax + 1
  __shared__ uint32_t warp_tot[8];
  __shared__ uint32_t out_n;
  __shared__ unsigned long long out_base;

  const uint32_t t = blockIdx.x;
  const unsigned long long off = J.tile_off[t];
  const uint32_t c = static_cast<uint32_t>(J.tile_off[t + 1] - off);
  if (c < 2 || c > J.cmax) return;                       // oversize tiles: k_tile_join_big
  const uint32_t tid = threadIdx.x;
  const uint32_t stride = J.stride;
  const uint32_t kshift = J.id_bits + 20;

  for (uint32_t b = tid; b <= kTjBuckets; b += 256) boff[b] = 0;
  if (tid == 0) out_n = 0;
  __syncthreads();
  unsigned long long e[3];
  uint32_t rank[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t i = tid + k * 256;
    e[k] = i < c ? J.entries[off + i] : 0ull;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (tid + k * 256 < c) rank[k] = atomicAdd(&boff[static_cast<uint32_t>(e[k] >> kshift) & (kTjBuckets - 1)], 1u);
  __syncthreads();
  tj_block_scan<4>(boff, kTjBuckets, warp_tot, tid);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (tid + k * 256 < c) ent[boff[static_cast<uint32_t>(e[k] >> kshift) & (kTjBuckets - 1)] + rank[k]] = e[k];
  __syncthreads();

  unsigned long long st_p = 0, st_s = 0, st_x = 0, st_r = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t s = tid + k * 256;
    if (s >= c) continue;
    const unsigned long long es = ent[s];
    const uint32_t b = static_cast<uint32_t>(es >> kshift) & (kTjBuckets - 1);
    const uint32_t lo = boff[b], hi = boff[b + 1];
    bool partner = false;
    for (uint32_t q = lo; q < hi && !partner; ++q) partner = q != s && tj_compatible(J, es, ent[q]);
    pref[s] = partner ? hi - s - 1 : 0u;
    if (partner) {
      if (STATS) st_r++;
      const uint64_t *w = J.words + static_cast<uint64_t>(tj_id(J, es)) * stride;
      uint64_t *r = rows + static_cast<size_t>(s) * stride;
      for (uint32_t x = 0; x < stride; ++x) r[x] = w[x];
    }
  }
  __syncthreads();
  tj_block_scan<3>(pref, c, warp_tot, tid);
  const uint32_t P = pref[c];

  const uint32_t K = J.K;
  const uint64_t kmask1 = K >= 64 ? ~0ull : (K > 32 ? (1ull << (2 * (K - 32))) - 1 : 0ull);
  const uint64_t kmask0 = K >= 32 ? ~0ull : (1ull << (2 * K)) - 1;
  const uint32_t lane = tid & 31u;
  for (uint32_t p0 = 0; p0 < P; p0 += 256) {             // uniform trip count: the link staging below is warp-collective
    const uint32_t p = p0 + tid;
    uint2 l0 = make_uint2(0, 0), l1 = make_uint2(0, 0);
    uint32_t nl = 0;
    if (p < P) {
      uint32_t lo = 0, hi = c;                           // last s with pref[s] <= p
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pref[mid] <= p) lo = mid; else hi = mid;
      }
      const uint32_t s = lo, q = s + 1 + (p - pref[s]);
      const unsigned long long es = ent[s], eq = ent[q];
      if (STATS) st_s++;
      if (tj_compatible(J, es, eq)) {
        if (STATS) st_p++;
        nl = tj_pair<STATS>(J, es, eq, rows + static_cast<size_t>(s) * stride, rows + static_cast<size_t>(q) * stride, kmask0, kmask1, st_x, l0, l1);
      }
    }
    // links of the warp -> the tile's stage: one shared-memory atomic per warp and iteration
    const uint32_t b1 = __ballot_sync(kFull, nl >= 1), b2 = __ballot_sync(kFull, nl >= 2);
    const uint32_t tot = __popc(b1) + __popc(b2);
    if (tot) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&out_n, tot);
      base = __shfl_sync(kFull, base, 0);
      const uint32_t lt = (1u << lane) - 1u;
      const uint32_t i0 = base + __popc(b1 & lt), i1 = base + __popc(b1) + __popc(b2 & lt);
      if (nl >= 1) {
        if (i0 < kTjOutCap) out[i0] = l0;
        else { const unsigned long long g = atomicAdd(J.edge_count, 1ull); if (g < J.edge_cap) J.edges[g] = l0; }   // stage full
      }
      if (nl >= 2) {
        if (i1 < kTjOutCap) out[i1] = l1;
        else { const unsigned long long g = atomicAdd(J.edge_count, 1ull); if (g < J.edge_cap) J.edges[g] = l1; }
      }
    }
  }
  __syncthreads();
  const uint32_t m = min(out_n, kTjOutCap);
  if (tid == 0 && m) out_base = atomicAdd(J.edge_count, static_cast<unsigned long long>(m));
  __syncthreads();
  if (m) {
    const unsigned long long base = out_base;
    for (uint32_t i = tid; i < m; i += 256)
      if (base + i < J.edge_cap) J.edges[base + i] = out[i];
  }
  if (STATS) {
    unsigned long long st_e = 0;
    for (uint32_t i = tid; i < c; i += 256) st_e++;
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) {
      st_e += __shfl_xor_sync(kFull, st_e, mm);
      st_p += __shfl_xor_sync(kFull, st_p, mm);
      st_s += __shfl_xor_sync(kFull, st_s, mm);
      st_x += __shfl_xor_sync(kFull, st_x, mm);
      st_r += __shfl_xor_sync(kFull, st_r, mm);
    }
    if ((tid & 31u) == 0) {
      atomicAdd(&J.stats[0], st_e);
      atomicAdd(&J.stats[1], st_p);
      atomicAdd(&J.stats[2], st_s);
      atomicAdd(&J.stats[3], st_x);
      atomicAdd(&J.stats[4], st_r);
    }
  }
}

// oversize tiles: every entry against every other entry of its tile, keys compared from a shared-memory
// chunk, sequences from global memory.  O(c^2) per tile, exact; only dense data reaches it.
template <bool STATS>
__global__ void __launch_bounds__(256) k_tile_join_big(TileJoinParams J) {
  __shared__ unsigned long long chunk[256];
  __shared__ PairStage stage[8];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  PairStage &S = stage[warp];
  uint32_t scnt = 0;
  const uint32_t nbig = *J.big_count;
  const uint32_t K = J.K;
  const uint64_t kmask1 = K >= 64 ? ~0ull : (K > 32 ? (1ull << (2 * (K - 32))) - 1 : 0ull);
  const uint64_t kmask0 = K >= 32 ? ~0ull : (1ull << (2 * K)) - 1;
  unsigned long long st_e = 0, st_p = 0, st_s = 0, st_x = 0;
  for (uint32_t b = 0; b < nbig; ++b) {
    const uint32_t t = J.big_tiles[b];
    const unsigned long long off = J.tile_off[t];
    const uint32_t c = static_cast<uint32_t>(J.tile_off[t + 1] - off);
    const uint32_t nblk = (c + 255u) / 256u;
    for (uint32_t ib = blockIdx.x; ib < nblk; ib += gridDim.x) {
      const uint32_t i = ib * 256u + tid;
      const bool valid = i < c;
      const unsigned long long e = valid ? J.entries[off + i] : 0ull;
      const uint32_t id = tj_id(J, e);
      const uint64_t *ri = J.words + static_cast<uint64_t>(id) * J.stride;
      if (STATS && valid) st_e++;
      for (uint32_t j0 = 0; j0 < c; j0 += 256u) {
        __syncthreads();
        chunk[tid] = (j0 + tid < c) ? J.entries[off + j0 + tid] : 0ull;
        __syncthreads();
        const uint32_t lim = min(256u, c - j0);
        for (uint32_t jj = 0; jj < lim; ++jj) {
          const unsigned long long f = chunk[jj];
          const bool hit = valid && tj_id(J, f) > id && tj_compatible(J, e, f);
          if (STATS && valid) st_s++;
          if (!__any_sync(kFull, hit)) continue;
          uint32_t mine = 0;
          uint2 l0 = make_uint2(0, 0), l1 = make_uint2(0, 0);
          if (hit) {
            if (STATS) st_p++;
            mine = tj_pair<STATS>(J, e, f, ri, J.words + static_cast<uint64_t>(tj_id(J, f)) * J.stride, kmask0, kmask1, st_x, l0, l1);
          }
          stage_push(S, scnt, mine, l0, l1, J.edges, J.edge_count, J.edge_cap, lane);
        }
      }
    }
  }
  stage_flush(S, scnt, J.edges, J.edge_count, J.edge_cap, lane);
  if (STATS) {
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) {
      st_e += __shfl_xor_sync(kFull, st_e, mm);
      st_p += __shfl_xor_sync(kFull, st_p, mm);
      st_s += __shfl_xor_sync(kFull, st_s, mm);
      st_x += __shfl_xor_sync(kFull, st_x, mm);
    }
    if (lane == 0 && (st_e | st_s)) {
      atomicAdd(&J.stats[0], st_e);
      atomicAdd(&J.stats[1], st_p);
      atomicAdd(&J.stats[2], st_s);
      atomicAdd(&J.stats[3], st_x);
    }
  }
}

}  // namespace swb
