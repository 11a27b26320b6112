// swarm_b200/csrc/d1_dist.cuh — multi-GPU clustering of ONE job (SURVEY.md §8 row e): vertex-partitioned label /
// generation relaxation with the exchange done by the kernel itself over NVLink peer memory.
//
// Single GPU (d1_kernels.cuh): key[v] = swarm<<32 | generation = min over links u->v of key[u]+1, to the fixed
// point; parent[v] = min { u : u->v, key[u]+1 == key[v] }  (closed form of src/algod1.cc:1185-1280, :673-718).
// Replicating that on every GPU does not scale: the link list of the whole job has to be gathered everywhere and
// every GPU relaxes all of it (measured, 2 x 10 M amplicons: 2.8 ms instead of 1.1).  Here every rank OWNS a
// block-cyclic share of the amplicons (blocks of 4096 ids: links point from abundant = low ids to rare = high ids,
// contiguous ranges would leave the first rank with most of the sources): their keys, parents and "lowered" bitmaps
// live only in its memory, and it relaxes only the links that leave its amplicons.  One persistent cooperative
// kernel per GPU runs, back to back:
//   route    the links this rank's join found are bucketed by owner(src) in shared memory and written straight into
//            the owners' inboxes (peer pointers; every inbox has one sub-region per sender, so space is reserved
//            with a LOCAL atomicAdd per CTA chunk and owner — the first cut used remote atomics and waited an
//            NVLink round trip per chunk — and runs are written coalesced);
//   rounds   out-links of vertices lowered in the previous round are relaxed: a local destination is an atomicMin,
//            a remote one becomes a 16-byte record (v, u, key[u]+1) appended to the owner's message log; after a
//            cross-GPU barrier every rank applies the records that arrived.  The barrier also ORs "did anybody
//            send or lower": all zero = fixed point;
//   parents  no second exchange: the log still holds every offer ever made, the owner keeps the smallest u
//            whose offer equals the final key.
// Cross-GPU barrier: every rank stores (epoch, flag) into its slot of every peer's flag array (system-scope
// release) and spins on its own array; a 5 s timeout turns a lost peer into an error instead of a hang.
// The symmetric buffers come from the caller (torch symmetric memory / CUDA IPC): plain pointers in the C ABI.
#pragma once
#include <cooperative_groups.h>
#include <cstddef>
#include "common.cuh"

namespace swb {

constexpr uint32_t kDistMaxWorld = 16;
constexpr int kDistU = 8;                         // items per thread and chunk
constexpr uint32_t kDistChunk = 256 * kDistU;
constexpr uint32_t kDistLogFactor = 4;            // message-log records per link slot of an inbox
constexpr uint32_t kDistBlock = 4096;             // ownership is block-cyclic: owner(v) = (v / kDistBlock) % world

struct __align__(16) DistRec {                    // one relaxation message; 16-byte aligned: moved with ONE 128-bit access
  uint32_t v, u;                                  // (as two 8-byte stores each record crossed NVLink as two half-filled sectors)
  unsigned long long cand;
};

// control block at the start of every rank's peer-visible buffer (same layout on all ranks).  Every inbox is split
// into one sub-region per SENDER, so a sender appends with a counter in its own memory (no remote atomics) and
// publishes the count together with its barrier flag.
struct DistCtl {
  unsigned long long bar[2][kDistMaxWorld];       // barrier slots: (epoch << 1) | flag, written by the peers
  unsigned long long links_cnt[kDistMaxWorld];    // [sender] links routed into this rank's inbox
  unsigned long long upd_cnt[kDistMaxWorld];      // [sender] records appended to this rank's message log so far
  unsigned int overflow;                          // this rank could not fit something into a peer's inbox
};
constexpr size_t kDistCtlBytes = 4096;

struct DistParams {
  uint32_t rank, world, n;
  const uint2 *edges;                             // links found by this rank's join
  uint64_t m_local;
  unsigned long long *key;                        // keys of the owned amplicons, by local index
  uint32_t *parent, *label, *generation;          // n each, by amplicon id; this rank fills the ids it owns
  uint32_t *bits;                                 // 3 * nwords bitmaps over the owned amplicons
  uint32_t nwords;
  uint32_t n_local;                               // owned amplicons, rounded up to whole blocks
  uint2 *my_links;                                // the inbox compacted: out-links of the owned amplicons
  unsigned long long *lcnt;                       // [2][kDistMaxWorld] append counters of THIS sender: links, messages
  unsigned char *peer[kDistMaxWorld];             // peer-visible buffer of every rank (peer[rank] = mine)
  uint64_t cap_links, cap_upd;                    // capacity (items) of one sender's sub-region
  unsigned long long epoch_base;                  // barrier epochs of this call start above it
  uint32_t dbg;                                   // profiling aid: bit 0 = skip the peer stores of the rounds (results are then wrong)
  unsigned long long *ts;                         // optional profiling aid: %globaltimer of thread 0 at phase boundaries (128 slots)
  unsigned int *gbar;                             // grid barrier words (SwGrid), zero before the first launch
  uint32_t *lflags;                               // 3 rotating "this rank sent or lowered something" words + [3] abort + [4] rounds + [5] barrier result
};

__device__ __forceinline__ uint32_t dist_owner(const DistParams &D, uint32_t v) { return (v / kDistBlock) % D.world; }
__device__ __forceinline__ uint32_t dist_local(const DistParams &D, uint32_t v) { return (v / kDistBlock) / D.world * kDistBlock + (v % kDistBlock); }
__device__ __forceinline__ uint32_t dist_global(const DistParams &D, uint32_t l) { return ((l / kDistBlock) * D.world + D.rank) * kDistBlock + (l % kDistBlock); }
__device__ __forceinline__ DistCtl *dist_ctl(const DistParams &D, uint32_t r) { return reinterpret_cast<DistCtl *>(D.peer[r]); }
// sub-region of sender `s` in the inboxes of rank `r`
__device__ __forceinline__ uint2 *dist_links(const DistParams &D, uint32_t r, uint32_t s) {
  return reinterpret_cast<uint2 *>(D.peer[r] + kDistCtlBytes) + static_cast<size_t>(s) * D.cap_links;
}
__device__ __forceinline__ DistRec *dist_upd(const DistParams &D, uint32_t r, uint32_t s) {
  return reinterpret_cast<DistRec *>(D.peer[r] + kDistCtlBytes + D.world * D.cap_links * sizeof(uint2)) + static_cast<size_t>(s) * D.cap_upd;
}

// Grid-wide barrier in global memory (count + generation).  cooperative_groups' grid.sync() needs a cooperative launch, and
// the driver never co-schedules two cooperative kernels on one GPU — which is exactly what a job whose ranks share a GPU
// (tests: several ranks on the single GPU of the box) needs.  With this barrier the kernel can be launched either way:
// cooperatively (one rank per GPU: co-residency of the grid guaranteed) or plainly with a grid small enough to be resident.
struct SwGrid {
  unsigned int *bar;                              // [0] arrivals [1] generation [2] set when a CTA waited longer than 5 s
  __device__ __forceinline__ void sync() const {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int gen = *reinterpret_cast<volatile unsigned int *>(&bar[1]);
      if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
        bar[0] = 0;
        __threadfence();
        atomicAdd(&bar[1], 1u);
      } else {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        for (uint32_t spins = 0; *reinterpret_cast<volatile unsigned int *>(&bar[1]) == gen; ++spins) {
          if ((spins & 1023u) == 1023u) {
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t1 - t0 > 5000000000ull) { bar[2] = 1u; break; }      // a CTA of this grid never became resident
          }
        }
      }
      __threadfence();
    }
    __syncthreads();
  }
};

struct DistSmem {
  uint32_t cnt[kDistMaxWorld + 1];
  uint32_t start[kDistMaxWorld + 1];
  unsigned long long base[kDistMaxWorld];
};

// CTA-wide: every thread contributes up to kDistU items with an owner (kNone = nothing); the items are counting-
// sorted by owner in shared memory and each owner's run is appended to this sender's sub-region of that owner's
// inbox: one LOCAL atomicAdd per owner and chunk reserves the space, consecutive threads write consecutive
// addresses over NVLink.
template <typename T, typename Region>
__device__ __forceinline__ void dist_scatter(const DistParams &D, DistSmem &sm, T *sorted, const T (&item)[kDistU], const uint32_t (&owner)[kDistU],
                                             Region region, unsigned long long *counters, uint64_t cap) {
  const uint32_t tid = threadIdx.x;
  if (tid <= kDistMaxWorld) sm.cnt[tid] = 0;
  __syncthreads();
  // rank of every item inside its owner's bucket: one shared-memory atomic per warp and owner (lanes with the same
  // owner are found with match.any; per-item atomics on 1-2 hot counters serialised 32 ways and cost ~10 us a chunk)
  uint32_t r[kDistU];
  const uint32_t lane = tid & 31u;
#pragma unroll
  for (int k = 0; k < kDistU; ++k) {
    const uint32_t peers = __match_any_sync(kFull, owner[k]);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader && owner[k] != kNone) base = atomicAdd(&sm.cnt[owner[k]], static_cast<uint32_t>(__popc(peers)));
    base = __shfl_sync(kFull, base, leader);
    r[k] = base + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t o = 0; o < D.world; ++o) { sm.start[o] = run; run += sm.cnt[o]; }
    sm.start[D.world] = run;
  }
  if (tid < D.world && sm.cnt[tid]) sm.base[tid] = atomicAdd(&counters[tid], static_cast<unsigned long long>(sm.cnt[tid]));
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kDistU; ++k)
    if (owner[k] != kNone) sorted[sm.start[owner[k]] + r[k]] = item[k];
  __syncthreads();
  const uint32_t total = sm.start[D.world];
  for (uint32_t i = tid; i < total; i += blockDim.x) {
    uint32_t o = 0;
    while (i >= sm.start[o + 1]) ++o;
    const unsigned long long dest = sm.base[o] + (i - sm.start[o]);
    if (dest >= cap) dist_ctl(D, D.rank)->overflow = 1u;
    else if (!(D.dbg & 1u) || sizeof(T) != sizeof(DistRec)) region(o)[dest] = sorted[i];
  }
  __syncthreads();
}

// All CTAs of all ranks.  Publishes this sender's append counters (`counters[p]` -> the [rank] entry of the array at
// `remote_offset` in peer p's control block), and returns the OR over the ranks of *lflag_word.
__device__ __forceinline__ uint32_t dist_barrier(const DistParams &D, const SwGrid &grid, unsigned long long epoch,
                                                 volatile uint32_t *lflag_word, unsigned long long *counters, size_t remote_offset) {
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    uint32_t f = 0;
    if (lane < D.world) {
      const unsigned long long val = (epoch << 1) | (*lflag_word ? 1ull : 0ull);
      if (counters) {
        *reinterpret_cast<volatile unsigned long long *>(D.peer[lane] + remote_offset + 8 * D.rank) =
            *reinterpret_cast<volatile unsigned long long *>(&counters[lane]);
      }
      __threadfence_system();
      volatile unsigned long long *slot = &dist_ctl(D, lane)->bar[epoch & 1][D.rank];
      *slot = val;
      __threadfence_system();
      volatile unsigned long long *mine = &dist_ctl(D, D.rank)->bar[epoch & 1][lane];
      unsigned long long t0, t1, got;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
      for (;;) {
        got = *mine;
        if ((got >> 1) == epoch) break;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 5000000000ull || reinterpret_cast<volatile uint32_t *>(D.lflags)[3]) {      // a peer never arrived
          reinterpret_cast<volatile uint32_t *>(D.lflags)[3] = 1;
          got = 0;
          break;
        }
      }
      f = static_cast<uint32_t>(got & 1ull);
    }
    f = __any_sync(kFull, f != 0) ? 1u : 0u;
    if (lane == 0) D.lflags[5] = f;
    __threadfence();
  }
  grid.sync();
  return reinterpret_cast<volatile uint32_t *>(D.lflags)[5];
}

__device__ __forceinline__ void dist_stamp(const DistParams &D, uint64_t tid, uint32_t &slot) {
  if (D.ts && tid == 0 && slot < 127) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    D.ts[slot++] = t;
  }
}

__global__ void __launch_bounds__(256, 3) k_cluster_dist(DistParams D) {
  const SwGrid grid{D.gbar};
  __shared__ DistSmem sm;
  __shared__ unsigned long long s_pref[kDistMaxWorld + 1];
  extern __shared__ __align__(16) unsigned char dist_dyn[];      // kDistChunk * 16 bytes: sorted staging
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  volatile uint32_t *lflags = D.lflags;
  DistCtl *me = dist_ctl(D, D.rank);
  unsigned long long epoch = D.epoch_base;
  uint32_t tslot = 0;
  dist_stamp(D, tid, tslot);

  // ---- init (local) + route the links this rank found to the owners of their sources
  for (uint64_t l = tid; l < D.n_local; l += nth) {
    const uint32_t v = dist_global(D, static_cast<uint32_t>(l));
    D.key[l] = static_cast<unsigned long long>(v) << 32;
    if (v < D.n) D.parent[v] = kNone;
  }
  for (uint64_t w = tid; w < 3ull * D.nwords; w += nth) D.bits[w] = 0;
  if (tid == 0) { lflags[0] = 0; lflags[1] = 0; lflags[2] = 0; lflags[3] = 0; }
  {
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
  }
  __threadfence_system();               // once per thread, not per chunk: a system fence waits for the NVLink acknowledgements
  if (D.ts) { grid.sync(); dist_stamp(D, tid, tslot); }           // init + route
  dist_barrier(D, grid, ++epoch, lflags + 0, D.lcnt, offsetof(DistCtl, links_cnt));
  dist_stamp(D, tid, tslot);                                      // barrier
  // compact the per-sender sub-regions of my link inbox into one list
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (uint32_t s = 0; s < D.world; ++s) {
      s_pref[s] = run;
      run += min(static_cast<unsigned long long>(D.cap_links), *reinterpret_cast<volatile unsigned long long *>(&me->links_cnt[s]));
    }
    s_pref[D.world] = run;
  }
  __syncthreads();
  const uint64_t m = s_pref[D.world];
  for (uint32_t s = 0; s < D.world; ++s) {
    const uint2 *src = dist_links(D, D.rank, s);
    const uint64_t cnt = s_pref[s + 1] - s_pref[s];
    for (uint64_t i = tid; i < cnt; i += nth) D.my_links[s_pref[s] + i] = __ldcg(&src[i]);
  }
  grid.sync();
  dist_stamp(D, tid, tslot);                                      // compaction
  const uint2 *links = D.my_links;

  // ---- rounds.  Messages are APPENDED to the receiver's log for the whole call (never overwritten): every round
  // applies the records that arrived since the previous barrier, and the parents are found at the end by
  // scanning the log once more, locally — no second exchange.
  DistRec *sorted = reinterpret_cast<DistRec *>(dist_dyn);
  unsigned long long *counters = D.lcnt + kDistMaxWorld;
  __shared__ unsigned long long s_prev[kDistMaxWorld], s_cur[kDistMaxWorld];
  if (threadIdx.x < kDistMaxWorld) { s_prev[threadIdx.x] = 0; s_cur[threadIdx.x] = 0; }
  __syncthreads();
  uint32_t round = 0;
  for (;; ++round) {
    const uint32_t *rd = D.bits + static_cast<size_t>(round % 3) * D.nwords;
    uint32_t *wr = D.bits + static_cast<size_t>((round + 1) % 3) * D.nwords;
    uint32_t *cl = D.bits + static_cast<size_t>((round + 2) % 3) * D.nwords;
    if (tid == 0) lflags[(round + 1) % 3] = 0;
    if (round) for (uint64_t w = tid; w < D.nwords; w += nth) cl[w] = 0;
    int ch = 0;
    bool scattered = false;                                    // CTA-uniform: some thread of this CTA wrote to a peer
    for (uint64_t base = static_cast<uint64_t>(blockIdx.x) * kDistChunk; base < m; base += static_cast<uint64_t>(gridDim.x) * kDistChunk) {
      uint2 ed[kDistU];
      uint32_t lu[kDistU];
      bool act[kDistU];
#pragma unroll
      for (int k = 0; k < kDistU; ++k) {
        const uint64_t i = base + static_cast<uint64_t>(k) * blockDim.x + threadIdx.x;
        act[k] = i < m;
        ed[k] = act[k] ? links[i] : make_uint2(0u, 0u);
        lu[k] = act[k] ? dist_local(D, ed[k].x) : 0u;
      }
      if (round) {
        uint32_t w[kDistU];
#pragma unroll
        for (int k = 0; k < kDistU; ++k) w[k] = act[k] ? rd[lu[k] >> 5] : 0u;
#pragma unroll
        for (int k = 0; k < kDistU; ++k) act[k] = act[k] && ((w[k] >> (lu[k] & 31u)) & 1u);
      }
      DistRec item[kDistU];
#pragma unroll
      for (int k = 0; k < kDistU; ++k) item[k].cand = (act[k] ? D.key[lu[k]] : ~0ull) + 1ull;
      uint32_t owner[kDistU], lv[kDistU];
      bool any_remote = false;
#pragma unroll
      for (int k = 0; k < kDistU; ++k) {
        item[k].v = ed[k].y; item[k].u = ed[k].x;
        const uint32_t o = act[k] ? dist_owner(D, ed[k].y) : kNone;
        lv[k] = o == D.rank ? dist_local(D, ed[k].y) : kNone;               // local destination
        owner[k] = (o != kNone && o != D.rank) ? o : kNone;                  // remote destination
        any_remote |= owner[k] != kNone;
      }
      unsigned long long kd[kDistU];
#pragma unroll
      for (int k = 0; k < kDistU; ++k) kd[k] = lv[k] != kNone ? D.key[lv[k]] : 0ull;      // all destination keys in flight together
#pragma unroll
      for (int k = 0; k < kDistU; ++k)
        if (lv[k] != kNone && item[k].cand < kd[k] && atomicMin(&D.key[lv[k]], item[k].cand) > item[k].cand) {
          atomicOr(&wr[lv[k] >> 5], 1u << (lv[k] & 31u));
          ch = 1;
        }
      if (__syncthreads_or(any_remote)) {
        ch |= any_remote;
        scattered = true;
        dist_scatter<DistRec>(D, sm, sorted, item, owner, [&](uint32_t o) { return dist_upd(D, o, D.rank); }, counters, D.cap_upd);
      }
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) lflags[round % 3] = 1;
    if (scattered) __threadfence_system();                     // this CTA's peer writes are visible system-wide before it arrives
    if (D.ts) { grid.sync(); dist_stamp(D, tid, tslot); }      // relax
    const uint32_t busy = dist_barrier(D, grid, ++epoch, lflags + (round % 3), counters, offsetof(DistCtl, upd_cnt));
    dist_stamp(D, tid, tslot);                                 // barrier
    if (!busy || lflags[3]) break;
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
          if (cand < kv[k] && atomicMin(&D.key[lv], cand) > cand) atomicOr(&wr[lv >> 5], 1u << (lv & 31u));
        }
      }
    }
    grid.sync();
    dist_stamp(D, tid, tslot);                                 // apply
  }

  // ---- parents: smallest u among the in-links that offer exactly the final key — local links, then the message log
  for (uint64_t base = tid; base < m; base += nth * kDistU) {
    uint2 ed[kDistU];
    unsigned long long ku[kDistU], kv[kDistU];
#pragma unroll
    for (int k = 0; k < kDistU; ++k) {
      const uint64_t i = base + static_cast<uint64_t>(k) * nth;
      ed[k] = i < m ? links[i] : make_uint2(kNone, kNone);
      if (ed[k].x != kNone && dist_owner(D, ed[k].y) != D.rank) ed[k].x = kNone;       // remote destination: its owner decides
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
  for (uint64_t l = tid; l < D.n_local; l += nth) {
    const uint32_t v = dist_global(D, static_cast<uint32_t>(l));
    if (v >= D.n) continue;
    const unsigned long long kv = D.key[l];
    D.label[v] = static_cast<uint32_t>(kv >> 32);
    D.generation[v] = static_cast<uint32_t>(kv);
  }
  if (tid == 0) lflags[4] = round + 1;
  // nobody starts the next call (and appends to an inbox) before every rank has finished reading its own
  dist_barrier(D, grid, ++epoch, lflags + 3, nullptr, 0);
  if (tid < 2 * kDistMaxWorld) D.lcnt[tid] = 0;            // this sender's counters restart with the next call
  dist_stamp(D, tid, tslot);                               // parents + unpack + last barrier
  if (D.ts && tid == 0) D.ts[tslot] = 0;
}

}  // namespace swb
