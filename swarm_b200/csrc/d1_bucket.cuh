// swarm_b200/csrc/d1_bucket.cuh — "Clustering" (src/algod1.cc:1185-1280, process_seed :673-718) with the links BUCKETED BY
// SOURCE BLOCK: one kernel for one GPU (world = 1) and for the multi-GPU job (the exchange over NVLink peer memory of
// d1_dist.cuh: same inboxes, same message records, same cross-GPU barrier).
//
// Closed form: key[v] = swarm << 32 | generation = min over links u -> v of key[u] + 1 to the fixed point;
// parent[v] = min { u : u -> v, key[u] + 1 == key[v] }.  What r1's kernels (k_cluster_persistent, k_cluster_dist) paid for,
// per ncu and per the %globaltimer stamps of profiles/r2i_*: every one of the ~15 rounds walked ALL links and tested one bit of a
// "lowered last round" bitmap per link at a random address — a 44 us floor per round however few amplicons were still moving
// (0.6 of 1.1 ms on one GPU, 1.6 of 2.5 ms in the multi-GPU kernel, whose per-chunk staging ran even for idle chunks) — and every
// ACTIVE link read key[src] at a random address.  Here, once, the links are counting-sorted by the 4096-id block of their
// source (a count pass of fire-and-forget atomics, one scan, one scatter pass: ~2 400 buckets at 10 M amplicons); then
//   * a round walks the bucketed list in units of 2 048 links and SKIPS every unit whose source blocks had no amplicon lowered
//     in the previous round (a warp ORs the 128 bitmap words of the unit's block): late rounds touch a few units, not 70 MB;
//   * inside a unit the sources span one block (or a few): their bitmap words (512 B) and keys (32 KB) are L1-resident, so only
//     key[dst] is a random access;
//   * round 0 is fused into the scatter pass (every key still has its initial value src << 32: no load);
//   * multi-GPU: a remote destination becomes a 16-byte record in the owner's message log; a unit reserves its range with one
//     atomicAdd per destination rank on the sender's LOCAL counters (ranks inside the unit from ballots + shared-memory
//     atomics) and writes straight from registers — no counting sort of the records in shared memory as in r1.
// Ownership is block-cyclic (blocks of kDistBlock = 4096 ids), so "source block" = ownership block: a bucket never mixes owners.
#pragma once
#include "d1_dist.cuh"

namespace swb {

constexpr uint32_t kBkUnit = 2048;                // links per work unit of a round (256 threads x 8)

struct BucketParams {
  DistParams D;                                   // rank, world, n, n_local, edges, m_local, key, parent, label, generation, bits, nwords, peers ...
  uint32_t nblk;                                  // owned blocks = n_local / kDistBlock
  uint32_t *bcount;                               // nblk: links per source block, then the scatter cursors
  unsigned long long *boff;                       // nblk + 1: first bucketed link of every block
  uint2 *blinks;                                  // the links, bucketed by source block
  uint64_t blinks_cap;
  uint2 *unit_blk;                                // per work unit: first and last source block of its links
  uint32_t *act_list;                             // units to walk in the current round
  uint32_t *act_n;                                // [2] rotating counters of act_list
  uint64_t unit_cap;
  uint32_t ib, gb;                                // PACK: bits of an amplicon id / of the generation in the packed word (d1_kernels.cuh)
};

// position of the first block whose links start after link p (boff is non-decreasing): block of link p
__device__ __forceinline__ uint32_t bk_block_of(const BucketParams &B, unsigned long long p) {
  uint32_t lo = 0, hi = B.nblk;                   // last b with boff[b] <= p
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (B.boff[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// CTA-collective: every thread holds U links; those with a remote destination (own[k] is another rank) append the record
// (v, u, cand) to the message log of that rank.  Ranks inside the CTA come from warp ballots + one shared-memory atomic per warp
// and destination, then ONE global atomicAdd per destination and call reserves the CTA's range on the sender's LOCAL counter
// (a first cut reserved per warp: 250 k atomics a round on the one counter of the peer, +170 us a round at 2 GPUs), and the
// lanes of a warp write consecutive 16-byte slots.
template <int U>
__device__ __forceinline__ bool bk_send_unit(const BucketParams &B, unsigned long long *counters, uint32_t *s_cnt, unsigned long long *s_base,
                                             const uint32_t (&own)[U], const uint2 (&ed)[U], const unsigned long long (&cand)[U], uint32_t lane) {
  const DistParams &D = B.D;
  bool any = false;
#pragma unroll
  for (int k = 0; k < U; ++k) any |= own[k] != kNone && own[k] != D.rank;
  if (!__syncthreads_or(any)) return false;
  if (threadIdx.x < kDistMaxWorld) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t off[U];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const bool remote = own[k] != kNone && own[k] != D.rank;
    off[k] = 0;
    if (!__any_sync(kFull, remote)) continue;
    for (uint32_t d = 0; d < D.world; ++d) {
      const uint32_t m = __ballot_sync(kFull, remote && own[k] == d);
      if (!m) continue;
      uint32_t base = 0;
      if (lane == static_cast<uint32_t>(__ffs(m) - 1)) base = atomicAdd(&s_cnt[d], static_cast<uint32_t>(__popc(m)));
      base = __shfl_sync(kFull, base, __ffs(m) - 1);
      if (remote && own[k] == d) off[k] = base + __popc(m & lt);
    }
  }
  __syncthreads();
  if (threadIdx.x < D.world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&counters[threadIdx.x], static_cast<unsigned long long>(s_cnt[threadIdx.x]));
  __syncthreads();
#pragma unroll
  for (int k = 0; k < U; ++k) {
    if (own[k] == kNone || own[k] == D.rank) continue;
    const unsigned long long slot = s_base[own[k]] + off[k];
    if (slot >= D.cap_upd) dist_ctl(D, D.rank)->overflow = 1u;
    else if (!(D.dbg & 1u)) {
      uint4 rec;
      rec.x = ed[k].y; rec.y = ed[k].x; rec.z = static_cast<uint32_t>(cand[k]); rec.w = static_cast<uint32_t>(cand[k] >> 32);
      *reinterpret_cast<uint4 *>(dist_upd(D, own[k], D.rank) + slot) = rec;
    }
  }
  return true;
}

// PACK: the relaxed word is swarm << (gb + ib) | generation << ib | parent and an offer is ((word[u] >> ib) + 1) << ib | u — the
// atomicMin that lowers the key settles the parent too (k_cluster_persistent in d1_kernels.cuh explains why that is the closed
// form), so the tail no longer re-reads the links and the whole message log to find the parents (0.26 of 2.9 ms at 8 GPUs).  A
// final generation of 2^gb - 1 or more raises lflags[6]; the last barrier ORs it over the ranks and the host runs the unpacked kernel.
template <bool PACK>
__global__ void __launch_bounds__(256, 3) k_cluster_bucket(BucketParams B) {
  const DistParams &D = B.D;
  const SwGrid grid{D.gbar};
  const uint32_t ib = B.ib, gb = B.gb;
  const unsigned long long idmask = PACK ? (1ull << ib) - 1ull : 0ull;
  const unsigned long long gmask = PACK ? (gb >= 32 ? 0xFFFFFFFFull : (1ull << gb) - 1ull) : 0ull;
  // offer of source u whose relaxed word is `ws`: its key part plus one, u in the parent field (a generation that wraps into the
  // swarm field is caught when the words are unpacked, see k_cluster_persistent)
  auto offer_of = [&](unsigned long long ws, uint32_t u) -> unsigned long long {
    return PACK ? (((ws | idmask) + 1ull) | u) : ws + 1ull;
  };
  // fire-and-forget atomicMin (RED.MIN) of an offer into the owned amplicon lv whose word was `seen` a moment ago (cand < seen);
  // true when its KEY is (being) lowered — decided from `seen`: a superset of the true set whose extra members were lowered, and
  // marked, by a concurrent better offer in this same round (k_cluster_persistent has the argument and the measurement)
  auto lower = [&](uint32_t lv, unsigned long long cand, unsigned long long seen) -> bool {
    atomicMin(&D.key[lv], cand);
    return PACK ? (seen | idmask) > (cand | idmask) : true;
  };
  __shared__ DistSmem sm;
  __shared__ unsigned long long s_pref[kDistMaxWorld + 1];
  __shared__ unsigned long long s_prev[kDistMaxWorld], s_cur[kDistMaxWorld];
  __shared__ unsigned long long scan_part[256];
  __shared__ uint32_t s_cnt[kDistMaxWorld];
  __shared__ unsigned long long s_base[kDistMaxWorld];
  extern __shared__ __align__(16) unsigned char dist_dyn[];      // kDistChunk * 16 bytes: staging of the link routing
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  volatile uint32_t *lflags = D.lflags;
  DistCtl *me = dist_ctl(D, D.rank);
  const bool multi = D.world > 1;
  unsigned long long epoch = D.epoch_base;
  uint32_t tslot = 0;
  dist_stamp(D, tid, tslot);

  // ---- init (local); multi-GPU: route the links this rank's join found to the owners of their sources
  for (uint64_t l = tid; l < D.n_local; l += nth) {
    const uint32_t v = dist_global(D, static_cast<uint32_t>(l));
    D.key[l] = PACK ? ((static_cast<unsigned long long>(v) << (gb + ib)) | idmask) : (static_cast<unsigned long long>(v) << 32);
    if (!PACK && v < D.n) D.parent[v] = kNone;
  }
  for (uint64_t w = tid; w < 3ull * D.nwords; w += nth) D.bits[w] = 0;
  for (uint64_t w = tid; w < B.nblk; w += nth) B.bcount[w] = 0;
  if (tid == 0) { lflags[0] = 0; lflags[1] = 0; lflags[2] = 0; lflags[3] = 0; lflags[6] = 0; }
  if (multi) {
    uint2 *sorted = reinterpret_cast<uint2 *>(dist_dyn);
    for (uint64_t base = static_cast<uint64_t>(blockIdx.x) * kDistChunk; base < D.m_local; base += static_cast<uint64_t>(gridDim.x) * kDistChunk) {
      uint2 item[kDistU];
      uint32_t owner[kDistU];
#pragma unroll
      for (int k = 0; k < kDistU; ++k) {
        const uint64_t i = base + static_cast<uint64_t>(k) * blockDim.x + threadIdx.x;
        if (i < D.m_local) { item[k] = D.edges[i]; owner[k] = dist_owner(D, item[k].x); }
        else { item[k] = make_uint2(0, 0); owner[k] = kNone; }
      }
      dist_scatter<uint2>(D, sm, sorted, item, owner, [&](uint32_t o) { return dist_links(D, o, D.rank); }, D.lcnt, D.cap_links);
    }
    __threadfence_system();
    dist_barrier(D, grid, ++epoch, lflags + 0, D.lcnt, offsetof(DistCtl, links_cnt));
  } else {
    grid.sync();
  }
  dist_stamp(D, tid, tslot);                                      // init + route + barrier
  // the links whose source this rank owns: the sender sub-regions of my inbox (multi) or the join's own list
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (uint32_t s = 0; s < D.world; ++s) {
      s_pref[s] = run;
      run += multi ? min(static_cast<unsigned long long>(D.cap_links), *reinterpret_cast<volatile unsigned long long *>(&me->links_cnt[s]))
                   : static_cast<unsigned long long>(D.m_local);
    }
    s_pref[D.world] = run;
  }
  if (threadIdx.x < kDistMaxWorld) { s_prev[threadIdx.x] = 0; s_cur[threadIdx.x] = 0; }
  __syncthreads();
  const uint64_t m = min(s_pref[D.world], static_cast<unsigned long long>(B.blinks_cap));
  auto link_at = [&](uint64_t i) -> uint2 {
    if (!multi) return __ldcs(&D.edges[i]);
    uint32_t s = 0;
    while (i >= s_pref[s + 1]) ++s;
    return __ldcg(&dist_links(D, D.rank, s)[i - s_pref[s]]);
  };

  // ---- bucket the links by the block of their source: count, scan, scatter (+ round 0)
  // (source blocks are very unevenly loaded — abundant amplicons have the links — so per-link global atomics pile up on a few
  // addresses: 0.5 ms for the count and 1.5 ms for the scatter in the first cut; histograms and ranks live in shared memory)
  uint32_t *hist = reinterpret_cast<uint32_t *>(dist_dyn);          // nblk counters when they fit the staging buffer
  const bool smem_hist = static_cast<size_t>(B.nblk) * 8 <= static_cast<size_t>(kDistChunk) * sizeof(DistRec);
  if (smem_hist) {
    for (uint32_t b = threadIdx.x; b < B.nblk; b += 256) hist[b] = 0;
    __syncthreads();
    for (uint64_t i = tid; i < m; i += nth) atomicAdd(&hist[dist_local(D, link_at(i).x) / kDistBlock], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < B.nblk; b += 256)
      if (hist[b]) atomicAdd(&B.bcount[b], hist[b]);
  } else {
    for (uint64_t i = tid; i < m; i += nth) atomicAdd(&B.bcount[dist_local(D, link_at(i).x) / kDistBlock], 1u);
  }
  grid.sync();
  if (blockIdx.x == 0) {                                           // exclusive scan of nblk counts by one CTA
    const uint32_t per = (B.nblk + 255u) / 256u;
    const uint32_t b0 = min(B.nblk, threadIdx.x * per), b1 = min(B.nblk, b0 + per);
    unsigned long long s = 0;
    for (uint32_t b = b0; b < b1; ++b) s += B.bcount[b];
    scan_part[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t off = 1; off < 256; off <<= 1) {
      const unsigned long long v = threadIdx.x >= off ? scan_part[threadIdx.x - off] : 0ull;
      __syncthreads();
      scan_part[threadIdx.x] += v;
      __syncthreads();
    }
    unsigned long long run = scan_part[threadIdx.x] - s;
    for (uint32_t b = b0; b < b1; ++b) {
      const uint32_t c = B.bcount[b];
      B.boff[b] = run;
      B.bcount[b] = 0;                                              // becomes the scatter cursor
      run += c;
    }
    if (threadIdx.x == 255) B.boff[B.nblk] = scan_part[255];
    if (threadIdx.x < 2) B.act_n[threadIdx.x] = 0;
  }
  grid.sync();
  // first and last source block of every work unit (a round then reads 8 bytes per unit instead of searching)
  const uint64_t n_units = min(static_cast<unsigned long long>((m + kBkUnit - 1) / kBkUnit), static_cast<unsigned long long>(B.unit_cap));
  for (uint64_t u = tid; u < n_units; u += nth) {
    const unsigned long long p0 = u * kBkUnit, p1 = min(static_cast<unsigned long long>(m), p0 + kBkUnit);
    B.unit_blk[u] = make_uint2(bk_block_of(B, p0), bk_block_of(B, p1 - 1));
  }
  dist_stamp(D, tid, tslot);                                      // count + scan
  unsigned long long *counters = D.lcnt + kDistMaxWorld;
  {
    uint32_t *wr = D.bits + D.nwords;                               // round 1 reads bitmap 1 / flags 1
    int ch = 0;
    // a CTA takes chunks of kBkUnit links: rank inside (chunk, block) from a shared-memory atomic, ONE global atomicAdd per
    // (chunk, block) reserves the range in the bucket
    uint32_t *base_s = hist + B.nblk;
    if (smem_hist) { __syncthreads(); for (uint32_t b = threadIdx.x; b < B.nblk; b += 256) hist[b] = 0; }
    for (uint64_t c0 = static_cast<uint64_t>(blockIdx.x) * kBkUnit; c0 < m; c0 += static_cast<uint64_t>(gridDim.x) * kBkUnit) {
      uint2 ed[kBkUnit / 256];
      uint32_t blk[kBkUnit / 256], rk[kBkUnit / 256];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < static_cast<int>(kBkUnit / 256); ++k) {
        const uint64_t i = c0 + static_cast<uint64_t>(k) * 256 + threadIdx.x;
        const bool in = i < m;
        ed[k] = in ? link_at(i) : make_uint2(kNone, kNone);
        blk[k] = in ? dist_local(D, ed[k].x) / kDistBlock : 0u;
        rk[k] = 0;
        if (in) rk[k] = smem_hist ? atomicAdd(&hist[blk[k]], 1u) : atomicAdd(&B.bcount[blk[k]], 1u);
      }
      if (smem_hist) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < static_cast<int>(kBkUnit / 256); ++k)
          if (ed[k].x != kNone && rk[k] == 0) base_s[blk[k]] = atomicAdd(&B.bcount[blk[k]], hist[blk[k]]);
        __syncthreads();
      }
      constexpr int U = static_cast<int>(kBkUnit / 256);
      uint32_t own[U], lv[U];
      unsigned long long kd[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const bool in = ed[k].x != kNone;
        if (in) B.blinks[B.boff[blk[k]] + (smem_hist ? base_s[blk[k]] : 0u) + rk[k]] = ed[k];
        own[k] = in ? (multi ? dist_owner(D, ed[k].y) : D.rank) : kNone;
        lv[k] = own[k] == D.rank ? dist_local(D, ed[k].y) : kNone;
      }
#pragma unroll
      for (int k = 0; k < U; ++k) kd[k] = lv[k] != kNone ? D.key[lv[k]] : 0ull;
      unsigned long long cand0[U];                                    // key[src] still has its initial value: swarm src, generation 0
#pragma unroll
      for (int k = 0; k < U; ++k)
        cand0[k] = offer_of(PACK ? ((static_cast<unsigned long long>(ed[k].x) << (gb + ib)) | idmask) : (static_cast<unsigned long long>(ed[k].x) << 32), ed[k].x);
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (lv[k] != kNone && cand0[k] < kd[k] && lower(lv[k], cand0[k], kd[k])) {
          atomicOr(&wr[lv[k] >> 5], 1u << (lv[k] & 31u));
          ch = 1;
        }
      }
      if (multi) {
        if (bk_send_unit<U>(B, counters, s_cnt, s_base, own, ed, cand0, lane)) ch = 1;
      }
      if (smem_hist) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < static_cast<int>(kBkUnit / 256); ++k)
          if (ed[k].x != kNone) hist[blk[k]] = 0;                   // only the touched counters
      }
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) lflags[0] = 1;
  }

  // ---- rounds
  uint32_t round = 0;
  for (;; ++round) {
    uint32_t *wr = D.bits + static_cast<size_t>((round + 1) % 3) * D.nwords;
    if (round) {
      const uint32_t *rd = D.bits + static_cast<size_t>(round % 3) * D.nwords;
      uint32_t *cl = D.bits + static_cast<size_t>((round + 2) % 3) * D.nwords;
      if (tid == 0) lflags[(round + 1) % 3] = 0;
      for (uint64_t w = tid; w < D.nwords; w += nth) cl[w] = 0;
      int ch = 0;
      // the units whose source blocks had an amplicon lowered in the previous round
      uint32_t *an = B.act_n + (round & 1u);
      // (a warp per unit ORs the 128 bitmap words of each of the unit's source blocks: no per-block flag array — a first cut
      // kept one flag byte per block and its 2.4 KB were hammered by every lowering: all of it lives in one or two L2 slices)
      constexpr uint32_t kWordsPerBlock = kDistBlock / 32;
      for (uint64_t u = tid >> 5; u < n_units; u += nth >> 5) {
        const uint2 bb = B.unit_blk[u];
        uint32_t any = 0;
        for (uint32_t b = bb.x; b <= bb.y && !any; ++b) {
          uint32_t acc = 0;
#pragma unroll
          for (uint32_t w = 0; w < kWordsPerBlock; w += 32) acc |= rd[static_cast<size_t>(b) * kWordsPerBlock + w + lane];
          any = __any_sync(kFull, acc != 0) ? 1u : 0u;
        }
        if (any && lane == 0) B.act_list[atomicAdd(an, 1u)] = static_cast<uint32_t>(u);
      }
      if (tid == 0) B.act_n[(round + 1) & 1u] = 0;
      grid.sync();
      const uint32_t n_act = *reinterpret_cast<volatile uint32_t *>(an);
      for (uint32_t a = blockIdx.x; a < n_act; a += gridDim.x) {
        const uint64_t unit = B.act_list[a];
        const unsigned long long p0 = unit * kBkUnit, p1 = min(static_cast<unsigned long long>(m), p0 + kBkUnit);
        // loads issued stage by stage (links, bitmap words, source keys, destination keys): a link is a chain of dependent
        // reads, and one link at a time leaves every warp with a single request in flight
        constexpr int U = static_cast<int>(kBkUnit / 256);
        uint2 ed[U];
        uint32_t lu[U], w[U];
        bool act[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const unsigned long long p = p0 + static_cast<unsigned long long>(k) * 256 + threadIdx.x;
          act[k] = p < p1;
          ed[k] = act[k] ? B.blinks[p] : make_uint2(0u, 0u);
          lu[k] = dist_local(D, ed[k].x);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) w[k] = act[k] ? rd[lu[k] >> 5] : 0u;
#pragma unroll
        for (int k = 0; k < U; ++k) act[k] = act[k] && ((w[k] >> (lu[k] & 31u)) & 1u);
        unsigned long long cand[U], kd[U];
        uint32_t own[U], lv[U];
#pragma unroll
        for (int k = 0; k < U; ++k) cand[k] = act[k] ? offer_of(D.key[lu[k]], ed[k].x) : ~0ull;
#pragma unroll
        for (int k = 0; k < U; ++k) {
          own[k] = act[k] ? (multi ? dist_owner(D, ed[k].y) : D.rank) : kNone;
          lv[k] = own[k] == D.rank ? dist_local(D, ed[k].y) : kNone;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) kd[k] = lv[k] != kNone ? D.key[lv[k]] : 0ull;
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (lv[k] != kNone && cand[k] < kd[k] && lower(lv[k], cand[k], kd[k])) {
            atomicOr(&wr[lv[k] >> 5], 1u << (lv[k] & 31u));
            ch = 1;
          }
        if (multi) {
          if (bk_send_unit<U>(B, counters, s_cnt, s_base, own, ed, cand, lane)) ch = 1;
        }
      }
      if (__syncthreads_or(ch) && threadIdx.x == 0) lflags[round % 3] = 1;
    }
    if (multi) __threadfence_system();                              // peer writes of this thread are visible system-wide before it arrives
    uint32_t busy;
    if (multi) busy = dist_barrier(D, grid, ++epoch, lflags + (round % 3), counters, offsetof(DistCtl, upd_cnt));
    else { grid.sync(); busy = lflags[round % 3]; }
    dist_stamp(D, tid, tslot);                                      // relax (+ barrier)
    if (!busy || lflags[3]) break;
    if (multi) {
      // apply what arrived in this round, one sub-region per sender
      if (threadIdx.x < D.world) {
        s_prev[threadIdx.x] = s_cur[threadIdx.x];
        s_cur[threadIdx.x] = min(static_cast<unsigned long long>(D.cap_upd), *reinterpret_cast<volatile unsigned long long *>(&me->upd_cnt[threadIdx.x]));
      }
      __syncthreads();
      for (uint32_t s = 0; s < D.world; ++s) {
        if (s == D.rank) continue;
        const DistRec *in = dist_upd(D, D.rank, s);
        for (uint64_t base = s_prev[s] + tid; base < s_cur[s]; base += nth * 4) {
          uint4 raw[4];
          unsigned long long kv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t i = base + static_cast<uint64_t>(k) * nth;
            raw[k] = i < s_cur[s] ? __ldcg(reinterpret_cast<const uint4 *>(in + i)) : make_uint4(kNone, 0, 0, 0);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) kv[k] = raw[k].x != kNone ? D.key[dist_local(D, raw[k].x)] : 0ull;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (raw[k].x == kNone) continue;
            const uint32_t lv = dist_local(D, raw[k].x);
            const unsigned long long cand = (static_cast<unsigned long long>(raw[k].w) << 32) | raw[k].z;
            if (cand < kv[k] && lower(lv, cand, kv[k])) {
              atomicOr(&wr[lv >> 5], 1u << (lv & 31u));
            }
          }
        }
      }
      grid.sync();
      dist_stamp(D, tid, tslot);                                    // apply
    }
  }

  // ---- parents (unpacked word only): smallest u among the in-links that offer exactly the final key — local links, then the message log
  for (uint64_t base = tid; !PACK && base < m; base += nth * kDistU) {
    uint2 ed[kDistU];
    unsigned long long ku[kDistU], kv[kDistU];
#pragma unroll
    for (int k = 0; k < kDistU; ++k) {
      const uint64_t i = base + static_cast<uint64_t>(k) * nth;
      ed[k] = i < m ? B.blinks[i] : make_uint2(kNone, kNone);
      if (multi && ed[k].x != kNone && dist_owner(D, ed[k].y) != D.rank) ed[k].x = kNone;       // remote destination: its owner decides
    }
#pragma unroll
    for (int k = 0; k < kDistU; ++k) {
      ku[k] = ed[k].x != kNone ? D.key[dist_local(D, ed[k].x)] : 0ull;
      kv[k] = ed[k].x != kNone ? D.key[dist_local(D, ed[k].y)] : 0ull;
    }
#pragma unroll
    for (int k = 0; k < kDistU; ++k)
      if (ed[k].x != kNone && ku[k] + 1ull == kv[k]) atomicMin(&D.parent[ed[k].y], ed[k].x);
  }
  if (multi && !PACK) {
    for (uint32_t s = 0; s < D.world; ++s) {
      if (s == D.rank) continue;
      const DistRec *in = dist_upd(D, D.rank, s);
      const uint64_t cnt = s_cur[s];
      for (uint64_t base = tid; base < cnt; base += nth * kDistU) {
        uint4 raw[kDistU];
        unsigned long long kv[kDistU];
#pragma unroll
        for (int k = 0; k < kDistU; ++k) {
          const uint64_t i = base + static_cast<uint64_t>(k) * nth;
          raw[k] = i < cnt ? __ldcg(reinterpret_cast<const uint4 *>(in + i)) : make_uint4(kNone, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < kDistU; ++k) kv[k] = raw[k].x != kNone ? D.key[dist_local(D, raw[k].x)] : 0ull;
#pragma unroll
        for (int k = 0; k < kDistU; ++k)
          if (raw[k].x != kNone && ((static_cast<unsigned long long>(raw[k].w) << 32) | raw[k].z) == kv[k]) atomicMin(&D.parent[raw[k].x], raw[k].y);
      }
    }
  }
  for (uint64_t l = tid; l < D.n_local; l += nth) {
    const uint32_t v = dist_global(D, static_cast<uint32_t>(l));
    if (v >= D.n) continue;
    const unsigned long long kv = D.key[l];
    if (PACK) {
      D.label[v] = static_cast<uint32_t>(kv >> (gb + ib));
      D.generation[v] = static_cast<uint32_t>((kv >> ib) & gmask);
      if (((kv >> ib) & gmask) == gmask) lflags[6] = 1;          // deeper than the packed word holds
      D.parent[v] = (kv & idmask) == idmask ? kNone : static_cast<uint32_t>(kv & idmask);
    } else {
      D.label[v] = static_cast<uint32_t>(kv >> 32);
      D.generation[v] = static_cast<uint32_t>(kv);
    }
  }
  if (tid == 0) lflags[4] = round + 1;
  if (PACK && !multi) grid.sync();                                 // lflags[6] complete before the host reads it
  if (multi) {
    // nobody starts the next call (and appends to an inbox) before every rank has finished reading its own; the barrier also
    // tells every rank whether ANY rank met a generation the packed word cannot hold (result in lflags[5])
    dist_barrier(D, grid, ++epoch, lflags + 6, nullptr, 0);
    if (tid < 2 * kDistMaxWorld) D.lcnt[tid] = 0;            // this sender's counters restart with the next call
  }
  dist_stamp(D, tid, tslot);                                 // parents + unpack (+ last barrier)
  if (D.ts && tid == 0) D.ts[tslot] = 0;
}

}  // namespace swb
