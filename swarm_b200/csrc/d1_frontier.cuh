// swarm_b200/csrc/d1_frontier.cuh — "Clustering" (src/algod1.cc:1185-1280, process_seed :673-718) as a FRONTIER relaxation
// over out-adjacency rows (cluster_kernel = 4; measured 1.76 ms at 10 M amplicons against 0.98 ms for the default k_cluster_persistent:
// building the rows costs 0.63 ms — DESIGN.md §3.4 has the comparison).
//
// Closed form (checked against the oracle's step-by-step greedy loop on every test case):
//   key[v] = swarm << 32 | generation = min over directed links u -> v of key[u] + 1, to the fixed point, key[v] <= v << 32;
//   parent[v] = min { u : u -> v, key[u] + 1 == key[v] }.
// r1's k_cluster_persistent (d1_kernels.cuh) walked the WHOLE link list in every one of the ~14 rounds and tested one bit of
// a "lowered last round" bitmap per link at a random address: ncu showed 1.8e8 random L2 sectors and a 44 us floor per round
// however few vertices were still moving.  Here a round costs its frontier only:
//   build    one pass over the link list (src, dst): the link goes into the fixed 8-slot out-row of its source (one atomicAdd
//            + one 4-byte store; longer rows spill into a short list) AND is relaxed at once — key[src] is still its initial
//            value src << 32, known without a load — which is round 0;
//   rounds   the frontier is a bitmap over the amplicons (n / 8 bytes).  A warp loads 32 words coalesced, skips the empty
//            ones, and lane l relaxes the out-row of vertex 32 w + l: degree, key and the 32-byte row are read at consecutive
//            addresses across the warp, only key[dst] is random.  Late rounds touch a few KB.  Three bitmaps rotate
//            (read / set / clear), one grid barrier per round;
//   parents  from the rows, in id order: parent[v] = atomicMin over the sources that offer exactly the final key.
// Everything is one cooperative launch; no host round trip.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace swb {

constexpr uint32_t kFrSlots = 8;          // out-links kept in a vertex's row (one 32-byte sector); the rest spill

struct FrontierParams {
  const uint2 *edges;                     // directed links (src, dst)
  uint64_t m;
  uint32_t n;
  unsigned long long *key;
  uint32_t *parent, *label, *generation;
  uint32_t *deg;                          // out-degree of every vertex (may exceed kFrSlots: the excess is in `spill`)
  uint32_t *adj;                          // n * kFrSlots
  uint2 *spill;
  unsigned long long *spill_n;
  uint64_t spill_cap;
  uint32_t *bits;                         // 3 rotating frontier bitmaps of nwords words
  uint32_t nwords;
  volatile uint32_t *flags;               // 3 rotating "somebody was lowered" words
  uint32_t *rounds_out;
  unsigned long long *ts;                 // optional: %globaltimer at phase boundaries (profiling aid)
};

__device__ __forceinline__ void fr_stamp(const FrontierParams &P, uint64_t tid, uint32_t &slot) {
  if (P.ts && tid == 0 && slot < 63) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    P.ts[slot++] = t;
  }
}

// offer cand to v; a lowered key puts v on the next frontier
__device__ __forceinline__ bool fr_offer(const FrontierParams &P, uint32_t *wr, uint32_t v, unsigned long long cand) {
  if (cand < P.key[v] && atomicMin(&P.key[v], cand) > cand) {
    atomicOr(&wr[v >> 5], 1u << (v & 31u));
    return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) k_cluster_frontier(FrontierParams P) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = tid >> 5, nwarps = nth >> 5;
  const uint32_t n = P.n;
  uint32_t tslot = 0;
  fr_stamp(P, tid, tslot);
  for (uint64_t v = tid; v < n; v += nth) {
    P.key[v] = static_cast<unsigned long long>(v) << 32;
    P.deg[v] = 0;
    __stcs(&P.parent[v], kNone);
  }
  for (uint64_t w = tid; w < 3ull * P.nwords; w += nth) P.bits[w] = 0;
  if (tid == 0) { P.flags[0] = 0; P.flags[1] = 0; P.flags[2] = 0; *P.spill_n = 0; }
  grid.sync();
  fr_stamp(P, tid, tslot);

  // ---- build the out-rows + round 0 (every vertex still holds its initial key)
  {
    uint32_t *wr = P.bits + P.nwords;                    // round 1 reads bitmap 1
    int ch = 0;
    for (uint64_t base = tid; base < P.m; base += nth * 4) {
      uint2 ed[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t e = base + static_cast<uint64_t>(k) * nth;
        ed[k] = e < P.m ? __ldcs(&P.edges[e]) : make_uint2(kNone, kNone);
      }
      uint32_t slot[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) slot[k] = ed[k].x != kNone ? atomicAdd(&P.deg[ed[k].x], 1u) : 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ed[k].x == kNone) continue;
        if (slot[k] < kFrSlots) P.adj[static_cast<uint64_t>(ed[k].x) * kFrSlots + slot[k]] = ed[k].y;
        else {
          const unsigned long long s = atomicAdd(P.spill_n, 1ull);
          if (s < P.spill_cap) P.spill[s] = ed[k];
        }
        ch |= fr_offer(P, wr, ed[k].y, (static_cast<unsigned long long>(ed[k].x) << 32) + 1ull) ? 1 : 0;
      }
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) P.flags[0] = 1;
  }
  grid.sync();
  fr_stamp(P, tid, tslot);
  const uint64_t n_spill = min(*P.spill_n, static_cast<unsigned long long>(P.spill_cap));

  uint32_t round = 0;
  if (P.flags[0]) {
    for (round = 1;; ++round) {
      const uint32_t *rd = P.bits + static_cast<size_t>(round % 3) * P.nwords;
      uint32_t *wr = P.bits + static_cast<size_t>((round + 1) % 3) * P.nwords;
      uint32_t *cl = P.bits + static_cast<size_t>((round + 2) % 3) * P.nwords;
      if (tid == 0) P.flags[(round + 1) % 3] = 0;
      for (uint64_t w = tid; w < P.nwords; w += nth) cl[w] = 0;
      int ch = 0;
      for (uint64_t w0 = warp * 32; w0 < P.nwords; w0 += nwarps * 32) {
        const uint64_t wi = w0 + lane;
        const uint32_t mine = wi < P.nwords ? rd[wi] : 0u;
        uint32_t nz = __ballot_sync(kFull, mine != 0);
        while (nz) {
          const uint32_t src = __ffs(nz) - 1;
          nz &= nz - 1;
          const uint32_t word = __shfl_sync(kFull, mine, src);
          const uint32_t u = static_cast<uint32_t>((w0 + src) << 5) + lane;
          const uint32_t d = ((word >> lane) & 1u) ? P.deg[u] : 0u;
          if (d) {
            const unsigned long long cand = P.key[u] + 1ull;
            const uint4 *row = reinterpret_cast<const uint4 *>(P.adj + static_cast<uint64_t>(u) * kFrSlots);
            const uint4 r0 = row[0];
            uint32_t nb[kFrSlots] = {r0.x, r0.y, r0.z, r0.w, 0, 0, 0, 0};
            if (d > 4) { const uint4 r1 = row[1]; nb[4] = r1.x; nb[5] = r1.y; nb[6] = r1.z; nb[7] = r1.w; }
            const uint32_t dd = min(d, kFrSlots);
            unsigned long long kd[kFrSlots];
#pragma unroll
            for (uint32_t k = 0; k < kFrSlots; ++k) kd[k] = k < dd ? P.key[nb[k]] : 0ull;      // all destination keys in flight together
#pragma unroll
            for (uint32_t k = 0; k < kFrSlots; ++k)
              if (k < dd && cand < kd[k] && atomicMin(&P.key[nb[k]], cand) > cand) {
                atomicOr(&wr[nb[k] >> 5], 1u << (nb[k] & 31u));
                ch = 1;
              }
          }
        }
      }
      // hubs: rows longer than kFrSlots keep their tail in the spill list, scanned whole (short unless the data is dense)
      for (uint64_t s = tid; s < n_spill; s += nth) {
        const uint2 ed = P.spill[s];
        if ((rd[ed.x >> 5] >> (ed.x & 31u)) & 1u) ch |= fr_offer(P, wr, ed.y, P.key[ed.x] + 1ull) ? 1 : 0;
      }
      if (__syncthreads_or(ch) && threadIdx.x == 0) P.flags[round % 3] = 1;
      grid.sync();
      fr_stamp(P, tid, tslot);
      if (P.flags[round % 3] == 0) break;
    }
  }

  // ---- parents + unpack
  for (uint64_t u = tid; u < n; u += nth) {
    const uint32_t d = P.deg[u];
    if (d == 0) continue;
    const unsigned long long cand = P.key[u] + 1ull;
    const uint4 *row = reinterpret_cast<const uint4 *>(P.adj + u * kFrSlots);
    const uint4 r0 = row[0];
    uint32_t nb[kFrSlots] = {r0.x, r0.y, r0.z, r0.w, 0, 0, 0, 0};
    if (d > 4) { const uint4 r1 = row[1]; nb[4] = r1.x; nb[5] = r1.y; nb[6] = r1.z; nb[7] = r1.w; }
    const uint32_t dd = min(d, kFrSlots);
    unsigned long long kd[kFrSlots];
#pragma unroll
    for (uint32_t k = 0; k < kFrSlots; ++k) kd[k] = k < dd ? P.key[nb[k]] : 0ull;
#pragma unroll
    for (uint32_t k = 0; k < kFrSlots; ++k)
      if (k < dd && cand == kd[k]) atomicMin(&P.parent[nb[k]], static_cast<uint32_t>(u));
  }
  for (uint64_t s = tid; s < n_spill; s += nth) {
    const uint2 ed = P.spill[s];
    if (P.key[ed.x] + 1ull == P.key[ed.y]) atomicMin(&P.parent[ed.y], ed.x);
  }
  for (uint64_t v = tid; v < n; v += nth) {
    const unsigned long long kv = P.key[v];
    __stcs(&P.label[v], static_cast<uint32_t>(kv >> 32));
    __stcs(&P.generation[v], static_cast<uint32_t>(kv));
  }
  if (tid == 0 && P.rounds_out) *P.rounds_out = round + 1;
  fr_stamp(P, tid, tslot);
  if (P.ts && tid == 0) P.ts[tslot] = 0;
}

}  // namespace swb
