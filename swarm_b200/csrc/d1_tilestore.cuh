// swarm_b200/csrc/d1_tilestore.cuh — the d=1 neighbour network as a partitioned pigeonhole join over a TILE STORE
// (enum_mode JOIN, default kernel since r2a).  Replaces the count / scan / scatter index + row-gathering join of
// d1_tilejoin.cuh, which r1's ncu captures showed to be bound by instruction issue and by 32-byte-sector over-fetch
// (2.0 GB of DRAM traffic for 0.9 GB of algorithmic bytes: 40-byte rows at a 40-byte stride, fetched word by word
// by one thread each; 0.93 G warp instructions, most of them the per-tile counting sort, two block scans, the
// partner search and a binary search per pair).
//
// Same mathematics (src/variants.cc:184-249 enumerates exactly the sequences at Levenshtein distance 1; one edit
// cannot touch both the first K and the last K nucleotides, so every linked pair shares a K-mer piece anchored at
// the start or at the end), different data movement:
//   k_ts_scatter   "Hashing sequences" (src/algod1.cc:1118-1139).  ONE pass: a CTA stages 256 consecutive packed rows in
//                  shared memory with one 1-D TMA bulk copy, every thread hashes the two pieces of its row and appends
//                  a FAT record — the 8-byte entry `key | abh7 | len13 | id` followed by the packed row — to the
//                  fixed-capacity slot of the tile the high hash bits select (one atomicAdd on an L2-resident cursor).
//                  No counting pass, no scan, no host round trip; a tile is one contiguous 16-byte aligned block.
//   k_ts_join      "Building network" (src/algod1.cc:558-670).  One CTA per tile: the whole tile arrives with ONE
//                  cp.async.bulk + mbarrier wait (SASS UBLKCP) — entries AND rows, every HBM byte is read exactly once,
//                  streaming; same-key entries are found through hash chains in shared memory (one atomicExch per
//                  entry: no sort, no scan); the pairs are queued and decided converged, one pair per thread, equal
//                  lengths by Hamming distance, lengths one apart by the shifted comparison (the lane-parallel
//                  check_variant, src/variants.cc:118-165).
//   k_ts_big       records that did not fit their tile's slot (a dense group sharing one K-mer) live in an overflow
//                  list and are compared against their tile and against each other from global memory — exact; the
//                  host falls back to the linear enumeration when the quadratic cost of that gets out of hand.
// Multi-GPU (BASELINE configs[4]: table sharded by hash range): a rank owns the tiles [t_lo, t_hi);
// k_ts_route hashes only the rank's OWN rows and writes each record — entry + row, so the packed sequences travel
// with their entries and no rank holds the whole database — into the owner's inbox over NVLink peer memory,
// k_ts_scatter_inbox files the arrivals into the local store.  See engine.cu: index_tilestore().
#pragma once
#include "d1_tilejoin.cuh"

namespace swb {

constexpr uint32_t kTsBuckets = 1024;     // hash-chain heads per tile
constexpr uint32_t kTsNil = 0xFFFFu;
constexpr uint32_t kTsRows = 256;         // rows one CTA of the scatter pass stages

struct TileStoreParams {
  const uint64_t *words;            // packed rows of this context: local row r = amplicon row_first + r
  const uint32_t *len;
  const uint64_t *abundance;        // per local row
  const uint64_t *ab_all;           // abundance of EVERY amplicon by id (single GPU / replicated database), or null ...
  const uint32_t *run_start;        // ... then equal abundances are recognised by the run table (sharded database)
  uint32_t n_runs;
  uint32_t n;                       // amplicons of the whole job
  uint32_t row_first, row_count;
  uint32_t stride, K, id_bits;
  uint64_t kmask0, kmask1;          // the first K nucleotides of a packed row: masks of words 0 and 1
  int sorted_desc, ncb;
  uint32_t n_tiles, t_lo, t_hi;     // global tile count; this context's tiles
  uint32_t q_cap, out_cap;          // join: pairs queued per pass, links staged per tile before one global atomicAdd
  uint32_t cap, rec_words;          // records per tile slot; words per record: 1 + stride (FAT: entry + packed row) or 1 (slim: the
                                    // join gathers the rows it needs from `words`, which must then hold every amplicon)
  const unsigned long long *row_base; // INDIRECT records (sharded database): a record is `entry, ref`; ref = word offset of the packed row from
                                    // row_base (the inbox the row arrived in), or bit 63 + word offset into `words` (this rank's own rows)
  unsigned long long *store;        // (t_hi - t_lo) * cap * rec_words
  uint32_t *cursor;                 // records appended per local tile (beyond cap: overflow list)
  unsigned long long *ovf;          // overflow records: `local tile | previous overflow record of the tile << 32`, then the record
  uint32_t *ovf_head;               // per local tile: its latest overflow record (kNone: none) — the records of a tile form a chain
  unsigned long long *ovf_count;    // [0] records in the overflow list [1] sum of their positions (~ pair tests / 2)
  uint64_t ovf_cap;
  uint64_t ovf_budget;              // k_ts_big gives up (sets *ovf_abort) when [1] exceeds it; 0 = never
  uint32_t *ovf_abort;
  uint2 *edges;
  unsigned long long *edge_count;
  uint64_t edge_cap;
  uint32_t *dup_flag;
  unsigned long long *appended;     // records this context appended to its tiles (counted when stats are collected)
  unsigned long long *stats;        // [0] entries joined [1] same-key pairs [2] chain steps [3] exact comparisons [4] rows staged
};

__device__ __forceinline__ const uint64_t *ts_u64(const unsigned long long *p) { return reinterpret_cast<const uint64_t *>(p); }
__device__ __forceinline__ unsigned long long ts_pack(const TileStoreParams &J, uint64_t h, uint32_t piece, uint32_t L, uint64_t ab, uint32_t id) {
  const uint32_t abh = static_cast<uint32_t>((ab * 0x9E3779B97F4A7C15ull) >> 57);
  const uint64_t key = (h << 1) | piece;               // truncated by the shift below
  return (key << (J.id_bits + 20)) | (static_cast<unsigned long long>(abh) << (J.id_bits + 13)) |
         (static_cast<unsigned long long>(L & 0x1FFFu) << J.id_bits) | id;
}
__device__ __forceinline__ uint32_t ts_id(const TileStoreParams &J, unsigned long long e) { return static_cast<uint32_t>(e & ((1ull << J.id_bits) - 1ull)); }
__device__ __forceinline__ uint32_t ts_len(const TileStoreParams &J, unsigned long long e) { return static_cast<uint32_t>(e >> J.id_bits) & 0x1FFFu; }
__device__ __forceinline__ uint32_t ts_abh(const TileStoreParams &J, unsigned long long e) { return static_cast<uint32_t>(e >> (J.id_bits + 13)) & 0x7Fu; }
__device__ __forceinline__ uint32_t ts_key(const TileStoreParams &J, unsigned long long e) { return static_cast<uint32_t>(e >> (J.id_bits + 20)); }
__device__ __forceinline__ bool ts_compatible(const TileStoreParams &J, unsigned long long e, unsigned long long f) {
  const uint32_t Le = ts_len(J, e), Lf = ts_len(J, f);
  return ts_key(J, e) == ts_key(J, f) && Le + 1 >= Lf && Lf + 1 >= Le;
}

constexpr unsigned long long kTsRefLocal = 1ull << 63;
// packed row of a slim / indirect record (rec[0] = entry, rec[1] = ref when indirect)
__device__ __forceinline__ const uint64_t *ts_row_of(const TileStoreParams &J, const unsigned long long *rec) {
  if (J.row_base) {
    const unsigned long long ref = rec[1];
    return (ref & kTsRefLocal) ? J.words + (ref & ~kTsRefLocal) : ts_u64(J.row_base + ref);
  }
  return J.words + static_cast<uint64_t>(ts_id(J, rec[0]) - J.row_first) * J.stride;
}

// append one fat record (entry + packed row) to local tile t, or to the overflow list when the slot is full
__device__ __forceinline__ void ts_append(const TileStoreParams &J, uint32_t t, unsigned long long e, const uint64_t *row) {
  const uint32_t pos = atomicAdd(&J.cursor[t], 1u);
  unsigned long long *dst;
  if (pos < J.cap) {
    dst = J.store + (static_cast<uint64_t>(t) * J.cap + pos) * J.rec_words;
  } else {
    const unsigned long long o = atomicAdd(&J.ovf_count[0], 1ull);
    atomicAdd(&J.ovf_count[1], static_cast<unsigned long long>(pos));
    if (o >= J.ovf_cap) return;                        // the host sees ovf_count[0] > ovf_cap and grows the list
    const uint32_t prev = atomicExch(&J.ovf_head[t], static_cast<uint32_t>(o));
    dst = J.ovf + o * (J.rec_words + 1);
    *dst++ = t | (static_cast<unsigned long long>(prev) << 32);
  }
  if ((J.rec_words & 1u) == 0 && pos < J.cap) {        // 16-byte aligned record: 128-bit stores
    ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst);
    d2[0] = make_ulonglong2(e, row[0]);
    for (uint32_t k = 1; k + 1 < J.rec_words; k += 2) d2[(k + 1) >> 1] = make_ulonglong2(row[k], row[k + 1]);
  } else {
    dst[0] = e;
    for (uint32_t k = 0; k + 1 < J.rec_words; ++k) dst[1 + k] = row[k];
  }
}

// stage the packed rows [r0, r0 + cnt) of this context in shared memory with one TMA bulk copy.  Ends with the data visible.
__device__ __forceinline__ void ts_stage_rows(const TileStoreParams &J, uint64_t *rows, uint64_t *bar, uint32_t r0, uint32_t cnt) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (cnt * J.stride * 8u + 15u) & ~15u;          // the row buffers end with >= 16 bytes of padding
    mbar_expect_tx(bar, bytes);
    tma_load_1d(rows, J.words + static_cast<uint64_t>(r0) * J.stride, bytes, bar);
  }
  mbar_wait(bar, 0);
}

// "Hashing sequences", single GPU: every local row -> two fat records in the local tiles
__global__ void __launch_bounds__(kTsRows) k_ts_scatter(TileStoreParams J) {
  extern __shared__ __align__(128) unsigned char ts_smem[];
  uint64_t *rows = reinterpret_cast<uint64_t *>(ts_smem);
  __shared__ uint64_t bar;
  const uint32_t r0 = blockIdx.x * kTsRows;
  const uint32_t cnt = min(kTsRows, J.row_count - r0);
  ts_stage_rows(J, rows, &bar, r0, cnt);
  uint32_t kept = 0;
  if (threadIdx.x < cnt) {
    const uint32_t r = r0 + threadIdx.x;
    const uint64_t *w = rows + static_cast<size_t>(threadIdx.x) * J.stride;
    const uint32_t L = J.len[r];
    const uint64_t ab = J.abundance[r];
#pragma unroll
    for (uint32_t piece = 0; piece < 2; ++piece) {
      const uint64_t h = piece_hash(w, J.stride, piece ? L - J.K : 0u, J.K, piece);
      const uint32_t tile = static_cast<uint32_t>(__umul64hi(h, static_cast<uint64_t>(J.n_tiles)));
      if (tile < J.t_lo || tile >= J.t_hi) continue;     // a pass over a replicated database: another rank's hash range
      ts_append(J, tile - J.t_lo, ts_pack(J, h, piece, L, ab, J.row_first + r), w);
      ++kept;
    }
  }
  if (J.stats) {
    kept = __reduce_add_sync(kFull, kept);
    if ((threadIdx.x & 31u) == 0 && kept) atomicAdd(J.appended, static_cast<unsigned long long>(kept));
  }
}

// equal abundances?  (the smaller id always links to the larger one in a database sorted by abundance; the reverse link
// exists only for equal abundances, src/algod1.cc:580-583)
__device__ __forceinline__ bool ts_same_abundance(const TileStoreParams &J, uint32_t a, uint32_t v) {
  if (J.ab_all) return J.ab_all[a] == J.ab_all[v];
  uint32_t lo = 0, hi = J.n_runs;                        // run of a: last run with run_start <= a; same abundance <=> v is inside it too
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (J.run_start[mid] <= a) lo = mid; else hi = mid;
  }
  return v >= J.run_start[lo] && v < J.run_start[lo + 1];
}

// links of one decided pair (cls: 0 identical, 1 one edit apart, 2 further).  A pair that shares BOTH pieces is owned by
// its prefix tile; a prefix-tile pair whose first K nucleotides differ (a collision of the truncated keys) is owned by its
// suffix tile — every linked pair is emitted exactly once.
template <bool STATS>
__device__ __forceinline__ uint32_t ts_links(const TileStoreParams &J, unsigned long long ei, unsigned long long ej, int cls, bool pfx_eq,
                                             unsigned long long &st_x, uint2 &l0, uint2 &l1) {
  if (((ts_key(J, ei) & 1u) != 0) == pfx_eq) return 0;
  if (STATS) st_x++;
  if (cls == 0) atomicExch(J.dup_flag, 1u);
  if (cls != 1) return 0;
  uint32_t a = ts_id(J, ei), v = ts_id(J, ej);
  if (J.ncb) { l0 = make_uint2(a, v); l1 = make_uint2(v, a); return 2; }
  if (J.sorted_desc) {
    if (a > v) { const uint32_t t_ = a; a = v; v = t_; }
    l0 = make_uint2(a, v);
    if (ts_abh(J, ei) == ts_abh(J, ej) && ts_same_abundance(J, a, v)) { l1 = make_uint2(v, a); return 2; }
    return 1;
  }
  const uint64_t aa = J.ab_all[a], av = J.ab_all[v];       // unsorted databases are single-GPU only (engine.cu)
  uint32_t nl = 0;
  if (aa >= av) { l0 = make_uint2(a, v); nl = 1; }
  if (av >= aa) { if (nl) l1 = make_uint2(v, a); else l0 = make_uint2(v, a); ++nl; }
  return nl;
}

// shared-memory atomic add issued by ONE lane the caller has already elected: plain `atomicAdd` makes the compiler wrap the
// instruction in its own warp aggregation (vote / find-leader / popc / shuffle, ~20 instructions) every time
__device__ __forceinline__ uint32_t atoms_add(uint32_t *p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add(uint32_t *p, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// equal lengths: 0 / 1 / 2+ differing positions; pfx_eq = the first K nucleotides agree.  One substitution changes one
// word: count the differing words (a select per word), popcount only the single survivor.
__device__ __forceinline__ int ts_classify_eq(const uint64_t *x, const uint64_t *y, uint32_t stride, uint64_t kmask0, uint64_t kmask1, bool &pfx_eq) {
  uint32_t nzw = 0;
  uint64_t d0 = 0, d1 = 0, dw = 0;
  for (uint32_t k = 0; k < stride; ++k) {
    const uint64_t a = x[k] ^ y[k];
    if (k == 0) d0 = a;
    if (k == 1) d1 = a;
    if (a != 0) { ++nzw; dw = a; }
  }
  pfx_eq = (d0 & kmask0) == 0 && (d1 & kmask1) == 0;
  if (nzw != 1) return nzw ? 2 : 0;
  return __popcll((dw | (dw >> 1)) & 0x5555555555555555ull) == 1 ? 1 : 2;
}

// warp-collective: lanes contribute nl (0..2) links; staged in the tile's shared-memory buffer, spilled to the global list
__device__ __forceinline__ void ts_stage_links(const TileStoreParams &J, uint2 *out, uint32_t *out_n, uint32_t nl, uint2 l0, uint2 l1, uint32_t lane) {
  const uint32_t kTsOut = J.out_cap;
  const uint32_t b1 = __ballot_sync(kFull, nl >= 1), b2 = __ballot_sync(kFull, nl >= 2);
  const uint32_t tot = __popc(b1) + __popc(b2);
  if (tot == 0) return;
  uint32_t base = 0;
  if (lane == 0) base = atoms_add(out_n, tot);
  base = __shfl_sync(kFull, base, 0);
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t i0 = base + __popc(b1 & lt), i1 = base + __popc(b1) + __popc(b2 & lt);
  if (nl >= 1) {
    if (i0 < kTsOut) out[i0] = l0;
    else { const unsigned long long g = atomicAdd(J.edge_count, 1ull); if (g < J.edge_cap) J.edges[g] = l0; }
  }
  if (nl >= 2) {
    if (i1 < kTsOut) out[i1] = l1;
    else { const unsigned long long g = atomicAdd(J.edge_count, 1ull); if (g < J.edge_cap) J.edges[g] = l1; }
  }
}

// asynchronous 8-byte global -> shared copy (SASS LDGSTS): the row gather of the slim join, no register staging
__device__ __forceinline__ void cp_async_8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// "Building network": one CTA per tile.
//  (1) the tile's records arrive with ONE TMA bulk copy (FAT: entry + packed row each; slim: 8-byte entries);
//  (2) counting sort of the record INDICES by the low 10 key bits (shared-memory atomics give the rank, one block scan the
//      offsets): same-key records become neighbours; a 32-bit descriptor `key19 | len13` per sorted position is all the
//      pairing needs;
//  (3) slim only: the packed rows of the records whose bucket run holds at least two records are requested from the
//      database now, asynchronously (LDGSTS), and arrive while the pairs are being enumerated;
//  (4) pairs: sorted position s is paired with the later positions of its bucket run.  The runs are short on average (2.3)
//      and long for dense groups (40+), so the enumeration is load-balanced inside each warp: the 32 run lengths of a
//      block of positions are prefix-summed with shuffles and lane l takes pair p0 + l, locating its source position by a
//      5-step shuffle search (r2a walked a chain per lane: one lane busy for 40 steps, ncu: 45 % of 0.9 G instructions).
//      Same-key, length-compatible pairs are queued — equal lengths from the bottom, lengths one apart from the top — with
//      one shared-memory atomic per warp; a full queue suspends the enumeration until its pairs have been decided;
//  (5) the pairs are decided converged, one per thread: differing-word count then one popcount for equal lengths, the
//      shifted comparison for lengths one apart (the lane-parallel check_variant, src/variants.cc:118-165).
template <bool FAT, bool STATS, int OCC>
__global__ void __launch_bounds__(256, OCC) k_ts_join(TileStoreParams J) {
  extern __shared__ __align__(128) unsigned char ts_smem[];
  const uint32_t rw = J.rec_words, stride = J.stride;
  unsigned long long *recs = reinterpret_cast<unsigned long long *>(ts_smem);          // cap * rec_words
  uint64_t *rows = reinterpret_cast<uint64_t *>(recs + static_cast<size_t>(J.cap) * rw);   // slim: cap * stride gathered rows
  uint32_t *boff = reinterpret_cast<uint32_t *>(rows + (FAT ? 0 : static_cast<size_t>(J.cap) * stride));   // kTsBuckets + 2
  uint32_t *sdesc = boff + kTsBuckets + 2;                                              // cap descriptors, by sorted position
  uint32_t *queue = sdesc + J.cap;                                                      // kTsQueue pairs: s | q << 16 (sorted positions)
  const uint32_t kTsQueue = J.q_cap, kTsOut = J.out_cap;
  uint2 *out = reinterpret_cast<uint2 *>(queue + kTsQueue);                             // kTsOut links
  uint16_t *order = reinterpret_cast<uint16_t *>(out + kTsOut);                         // cap: record index of every sorted position
  __shared__ uint64_t bar;
  __shared__ uint32_t qn, qok, out_n;
  __shared__ unsigned long long out_base;
  __shared__ uint32_t warp_tot[8];

  const uint32_t t = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t c = min(J.cursor[t], J.cap);
  if (c < 2) return;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    qn = 0;
    qok = 0;
    out_n = 0;
  }
  for (uint32_t b = tid; b < kTsBuckets + 2; b += 256) boff[b] = 0;
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (c * rw * 8u + 15u) & ~15u;
    mbar_expect_tx(&bar, bytes);
    tma_load_1d(recs, J.store + static_cast<uint64_t>(t) * J.cap * rw, bytes, &bar);
  }
  mbar_wait(&bar, 0);

  const uint32_t kshift = J.id_bits + 20;
  uint32_t rank[3], desc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t i = tid + k * 256;
    if (i < c) {
      const unsigned long long e = recs[static_cast<size_t>(i) * rw];
      desc[k] = (static_cast<uint32_t>(e >> kshift) << 13) | ts_len(J, e);             // key bits 0..18 | length
      rank[k] = atoms_add(&boff[(desc[k] >> 13) & (kTsBuckets - 1)], 1u);      // raw ATOMS: the compiler's aggregation loop was 8.9 % of the kernel's instructions (r2d)
    }
  }
  __syncthreads();
  tj_block_scan<4>(boff, kTsBuckets, warp_tot, tid);            // boff[b] = first sorted position of bucket b, boff[kTsBuckets] = c
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t i = tid + k * 256;
    if (i < c) {
      const uint32_t pos = boff[(desc[k] >> 13) & (kTsBuckets - 1)] + rank[k];
      order[pos] = static_cast<uint16_t>(i);
      sdesc[pos] = desc[k];
    }
  }
  __syncthreads();
  unsigned long long st_p = 0, st_s = 0, st_x = 0, st_r = 0;
  if (!FAT) {                                                     // rows of the records that have a bucket mate: on their way
    for (uint32_t sp = tid; sp < c; sp += 256) {
      const uint32_t b = (sdesc[sp] >> 13) & (kTsBuckets - 1);
      if (boff[b + 1] - boff[b] >= 2u) {
        const uint32_t i = order[sp];
        const uint64_t *w = ts_row_of(J, recs + static_cast<size_t>(i) * rw);
        uint64_t *r = rows + static_cast<size_t>(i) * stride;
        for (uint32_t x = 0; x < stride; ++x) cp_async_8(r + x, w + x);
        if (STATS) st_r++;
      }
    }
  }

  const uint64_t kmask0 = J.kmask0, kmask1 = J.kmask1;
  const uint32_t nblk = (c + 31u) >> 5, lt = (1u << lane) - 1u;
  uint32_t blk = warp, p0 = 0;                                    // resumable: block of 32 sorted positions, first pair of the step
  for (bool first = true;; first = false) {
    bool fail = false;
    while (blk < nblk && !fail) {
      const uint32_t sp = (blk << 5) + lane;
      const uint32_t ds = sp < c ? sdesc[sp] : 0u;
      const uint32_t cnt = sp < c ? boff[((ds >> 13) & (kTsBuckets - 1)) + 1] - sp - 1u : 0u;
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
      }
      const uint32_t T = __shfl_sync(kFull, incl, 31);
      for (; p0 < T; p0 += 32) {
        const uint32_t p = p0 + lane;
        uint32_t lo = 0;                                          // lanes whose pairs all come before p
#pragma unroll
        for (uint32_t step = 16; step >= 1; step >>= 1) {
          const uint32_t v = __shfl_sync(kFull, incl, (lo + step - 1) & 31u);
          if (v <= p) lo += step;
        }
        const uint32_t src = lo & 31u;
        const uint32_t d_src = __shfl_sync(kFull, ds, src), excl_src = __shfl_sync(kFull, incl - cnt, src);
        const uint32_t ss = (blk << 5) + src, q = ss + 1u + (p - excl_src);
        bool ok = p < T;
        const uint32_t dq = ok ? sdesc[q] : 0u;
        const uint32_t ls = d_src & 0x1FFFu, lq = dq & 0x1FFFu;
        if (STATS && ok) st_s++;
        ok = ok && ((d_src ^ dq) >> 13) == 0 && ls + 1u >= lq && lq + 1u >= ls;
        const bool ne = ls != lq;
        const uint32_t m_eq = __ballot_sync(kFull, ok && !ne), m_ne = __ballot_sync(kFull, ok && ne);
        if (m_eq | m_ne) {
          const uint32_t k_eq = __popc(m_eq), k_ne = __popc(m_ne);
          uint32_t old = 0;
          if (lane == 0) old = atoms_add(&qn, k_eq | (k_ne << 16));
          old = __shfl_sync(kFull, old, 0);
          const uint32_t o_eq = old & 0xFFFFu, o_ne = old >> 16;
          if (o_eq + o_ne + k_eq + k_ne > kTsQueue) { fail = true; break; }       // full: this step is retried in the next pass
          if (lane == 0) reds_add(&qok, k_eq | (k_ne << 16));
          if (ok && !ne) queue[o_eq + __popc(m_eq & lt)] = ss | (q << 16);
          if (ok && ne) queue[kTsQueue - 1u - o_ne - __popc(m_ne & lt)] = ss | (q << 16);
          if (STATS && ok) st_p++;
        }
      }
      if (!fail) { blk += 8; p0 = 0; }
    }
    const bool more = __syncthreads_or(fail);
    if (!FAT && first) {
      cp_async_wait_all();
      __syncthreads();
    }
    // every push before the first failure wrote its slots: equal lengths [0, n_eq), lengths one apart (top - n_ne, top]
    const uint32_t n_eq = qok & 0xFFFFu, n_ne = qok >> 16;
    for (uint32_t q0 = 0; q0 < n_eq; q0 += 256) {
      const uint32_t p = q0 + tid;
      uint2 l0 = make_uint2(0, 0), l1 = make_uint2(0, 0);
      uint32_t nl = 0;
      if (p < n_eq) {
        const uint32_t pr = queue[p], a = order[pr & 0xFFFFu], b = order[pr >> 16];
        const unsigned long long ea = recs[static_cast<size_t>(a) * rw], eb = recs[static_cast<size_t>(b) * rw];
        const uint64_t *ra = FAT ? ts_u64(recs + static_cast<size_t>(a) * rw + 1) : rows + static_cast<size_t>(a) * stride;
        const uint64_t *rb = FAT ? ts_u64(recs + static_cast<size_t>(b) * rw + 1) : rows + static_cast<size_t>(b) * stride;
        bool pfx_eq;
        const int cls = ts_classify_eq(ra, rb, stride, kmask0, kmask1, pfx_eq);
        nl = ts_links<STATS>(J, ea, eb, cls, pfx_eq, st_x, l0, l1);
      }
      ts_stage_links(J, out, &out_n, nl, l0, l1, lane);
    }
    for (uint32_t q0 = 0; q0 < n_ne; q0 += 256) {
      const uint32_t p = q0 + tid;
      uint2 l0 = make_uint2(0, 0), l1 = make_uint2(0, 0);
      uint32_t nl = 0;
      if (p < n_ne) {
        const uint32_t pr = queue[kTsQueue - 1u - p], a = order[pr & 0xFFFFu], b = order[pr >> 16];
        const unsigned long long ea = recs[static_cast<size_t>(a) * rw], eb = recs[static_cast<size_t>(b) * rw];
        const uint64_t *ra = FAT ? ts_u64(recs + static_cast<size_t>(a) * rw + 1) : rows + static_cast<size_t>(a) * stride;
        const uint64_t *rb = FAT ? ts_u64(recs + static_cast<size_t>(b) * rw + 1) : rows + static_cast<size_t>(b) * stride;
        bool pfx_eq;
        const int cls = tj_classify(ra, ts_len(J, ea), rb, ts_len(J, eb), stride, kmask0, kmask1, pfx_eq);
        nl = ts_links<STATS>(J, ea, eb, cls, pfx_eq, st_x, l0, l1);
      }
      ts_stage_links(J, out, &out_n, nl, l0, l1, lane);
    }
    __syncthreads();
    if (!more) break;
    if (tid == 0) { qn = 0; qok = 0; }
    __syncthreads();
  }
  const uint32_t m = min(out_n, kTsOut);
  if (tid == 0 && m) out_base = atomicAdd(J.edge_count, static_cast<unsigned long long>(m));
  __syncthreads();
  if (m) {
    const unsigned long long base = out_base;
    for (uint32_t k = tid; k < m; k += 256)
      if (base + k < J.edge_cap) J.edges[base + k] = out[k];
  }
  if (STATS) {
    unsigned long long st_e = 0;
    for (uint32_t k = tid; k < c; k += 256) st_e++;
    if (FAT) st_r = st_e;
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) {
      st_e += __shfl_xor_sync(kFull, st_e, mm);
      st_p += __shfl_xor_sync(kFull, st_p, mm);
      st_s += __shfl_xor_sync(kFull, st_s, mm);
      st_x += __shfl_xor_sync(kFull, st_x, mm);
      st_r += __shfl_xor_sync(kFull, st_r, mm);
    }
    if (lane == 0) {
      atomicAdd(&J.stats[0], st_e);
      atomicAdd(&J.stats[1], st_p);
      atomicAdd(&J.stats[2], st_s);
      atomicAdd(&J.stats[3], st_x);
      atomicAdd(&J.stats[4], st_r);
    }
  }
}

// overflow records: against the records in their tile's slot (one warp per record, lanes over the slot), then against the
// earlier overflow records of the same tile (one thread per record, down the tile's chain).  Rows are read from global
// memory.  Exact; O(overflow x records of the tile).
template <bool FAT, bool STATS>
__global__ void __launch_bounds__(256) k_ts_big(TileStoreParams J) {
  __shared__ PairStage stage[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  PairStage &S = stage[warp];
  uint32_t scnt = 0;
  const unsigned long long n_ovf = min(J.ovf_count[0], static_cast<unsigned long long>(J.ovf_cap));
  if (n_ovf == 0) return;
  if (J.ovf_budget && J.ovf_count[1] > J.ovf_budget) {           // quadratic cost out of hand: the host takes another route
    if (blockIdx.x == 0 && threadIdx.x == 0) *J.ovf_abort = 1u;
    return;
  }
  const uint32_t rw = J.rec_words;
  const uint64_t kmask0 = J.kmask0, kmask1 = J.kmask1;
  unsigned long long st_s = 0, st_p = 0, st_x = 0;
  const uint64_t nwarps = static_cast<uint64_t>(gridDim.x) * 8;
  for (uint64_t x = static_cast<uint64_t>(blockIdx.x) * 8 + warp; x < n_ovf; x += nwarps) {
    const unsigned long long *ox = J.ovf + x * (rw + 1);
    const uint32_t t = static_cast<uint32_t>(ox[0]);
    const unsigned long long ex = ox[1];
    const unsigned long long *tile = J.store + static_cast<uint64_t>(t) * J.cap * rw;
    for (uint32_t q0 = 0; q0 < J.cap; q0 += 32) {
      const uint32_t q = q0 + lane;
      uint32_t mine = 0;
      uint2 l0 = make_uint2(0, 0), l1 = make_uint2(0, 0);
      if (q < J.cap) {
        const unsigned long long *rec = tile + static_cast<uint64_t>(q) * rw;
        if (STATS) st_s++;
        if (ts_compatible(J, ex, rec[0])) {
          if (STATS) st_p++;
          bool pfx_eq;
          const uint64_t *rx = FAT ? ts_u64(ox + 2) : ts_row_of(J, ox + 1);
          const uint64_t *ry = FAT ? ts_u64(rec + 1) : ts_row_of(J, rec);
          const int cls = tj_classify(rx, ts_len(J, ex), ry, ts_len(J, rec[0]), J.stride, kmask0, kmask1, pfx_eq);
          mine = ts_links<STATS>(J, ex, rec[0], cls, pfx_eq, st_x, l0, l1);
        }
      }
      stage_push(S, scnt, mine, l0, l1, J.edges, J.edge_count, J.edge_cap, lane);
    }
  }
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t x0 = static_cast<uint64_t>(blockIdx.x) * blockDim.x; x0 < n_ovf; x0 += nth) {
    const uint64_t x = x0 + threadIdx.x;
    const unsigned long long *ox = J.ovf + (x < n_ovf ? x : 0) * (rw + 1);
    const unsigned long long ex = ox[1];
    uint32_t y = x < n_ovf ? static_cast<uint32_t>(ox[0] >> 32) : kNone;
    while (__any_sync(kFull, y != kNone)) {
      uint32_t mine = 0;
      uint2 l0 = make_uint2(0, 0), l1 = make_uint2(0, 0);
      if (y != kNone) {
        const unsigned long long *oy = J.ovf + static_cast<uint64_t>(y) * (rw + 1);
        if (STATS) st_s++;
        if (ts_compatible(J, ex, oy[1])) {
          if (STATS) st_p++;
          bool pfx_eq;
          const uint64_t *rx = FAT ? ts_u64(ox + 2) : ts_row_of(J, ox + 1);
          const uint64_t *ry = FAT ? ts_u64(oy + 2) : ts_row_of(J, oy + 1);
          const int cls = tj_classify(rx, ts_len(J, ex), ry, ts_len(J, oy[1]), J.stride, kmask0, kmask1, pfx_eq);
          mine = ts_links<STATS>(J, ex, oy[1], cls, pfx_eq, st_x, l0, l1);
        }
        y = static_cast<uint32_t>(oy[0] >> 32);
      }
      stage_push(S, scnt, mine, l0, l1, J.edges, J.edge_count, J.edge_cap, lane);
    }
  }
  stage_flush(S, scnt, J.edges, J.edge_count, J.edge_cap, lane);
  if (STATS) {
#pragma unroll
    for (int mm = 16; mm >= 1; mm >>= 1) {
      st_p += __shfl_xor_sync(kFull, st_p, mm);
      st_s += __shfl_xor_sync(kFull, st_s, mm);
      st_x += __shfl_xor_sync(kFull, st_x, mm);
    }
    if (lane == 0 && st_s) {
      atomicAdd(&J.stats[1], st_p);
      atomicAdd(&J.stats[2], st_s);
      atomicAdd(&J.stats[3], st_x);
    }
  }
}

}  // namespace swb
