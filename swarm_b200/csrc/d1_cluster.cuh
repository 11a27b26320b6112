// swarm_b200/csrc/d1_cluster.cuh — clustering as ONE persistent cooperative kernel over
// links sorted by source.
//
// Closed form of the reference's greedy loop (src/algod1.cc:1185-1280, process_seed :673-718; d1_kernels.cuh has
// the derivation): key[v] = swarm<<32 | generation = min over links u->v of key[u]+1, iterated to the fixed
// point; parent[v] = min { u : u->v, key[u]+1 == key[v] }.
//
// What bounded the earlier kernels was not the arithmetic but random 32-byte L2 sectors: every round touched
// key[src], key[dst] (or a "changed" bit of src) of ALL links, in link-list order, i.e. at random —
// ~2 x 10^8 sectors, 1.1 ms at 10 M amplicons whatever was skipped afterwards.  Here the kernel first
// counting-sorts the links by source (degree histogram, grid-wide exclusive scan, fill): after that the source
// side of a round (source id, its changed bit, its key) is read in id order — coalesced and almost free — and
// only the destinations of links whose source was lowered in the previous round are touched at random.
// The active set decays geometrically with the BFS depth, so all rounds together cost ~2.6 passes.
// Phases are separated by grid-wide barriers (cooperative launch); the host launches once and never looks.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace swb {

struct ClusterParams {
  const uint2 *edges;            // directed links (src, dst), any order
  uint64_t m;
  uint32_t n;
  unsigned long long *key;       // n
  uint32_t *parent, *label, *generation;   // n each
  uint32_t *deg;                 // n: out-degree, then fill countdown
  uint32_t *row;                 // n + 1: exclusive prefix of deg
  uint32_t *srcs, *dsts;         // m each: the links sorted by source
  uint32_t *bits;                // 3 * nwords: rotating "lowered in round r" bitmaps
  uint32_t nwords;
  unsigned long long *cta_tot;   // gridDim.x
  volatile uint32_t *flags;      // 3
  uint32_t *rounds_out;
  unsigned long long *ts;        // optional: %globaltimer of thread 0 at every phase boundary (profiling aid), 64 slots
};

__device__ __forceinline__ void cl_stamp(const ClusterParams &C, uint64_t tid, uint32_t &slot) {
  if (C.ts && tid == 0 && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    C.ts[slot++] = t;
  }
}

constexpr int kCsU = 8;          // links per thread and step: loads are issued stage by stage to keep 8 requests in flight

__global__ void __launch_bounds__(256) k_cluster_csr(ClusterParams C) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ unsigned long long warp_tot[8];
  __shared__ unsigned long long cta_base;
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t n = C.n;
  const uint64_t m = C.m;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t slot = 0;
  cl_stamp(C, tid, slot);

  // ---- init
  for (uint64_t v = tid; v < n; v += nth) { C.key[v] = static_cast<unsigned long long>(v) << 32; C.parent[v] = kNone; C.deg[v] = 0; }
  for (uint64_t w = tid; w < 3ull * C.nwords; w += nth) C.bits[w] = 0;
  if (tid == 0) { C.flags[0] = 0; C.flags[1] = 0; C.flags[2] = 0; }
  grid.sync();
  cl_stamp(C, tid, slot);
  // ---- out-degrees
  for (uint64_t e = tid; e < m; e += nth) atomicAdd(&C.deg[C.edges[e].x], 1u);
  grid.sync();
  cl_stamp(C, tid, slot);
  // ---- exclusive scan of deg -> row: CTA b owns ids [b*chunk, (b+1)*chunk), thread t a run of `per` of them
  const uint64_t chunk = (static_cast<uint64_t>(n) + gridDim.x - 1) / gridDim.x;
  const uint64_t per = (chunk + 255) / 256;
  const uint64_t c_lo = min(static_cast<uint64_t>(n), blockIdx.x * chunk), c_hi = min(static_cast<uint64_t>(n), c_lo + chunk);
  const uint64_t t_lo = min(c_hi, c_lo + threadIdx.x * per), t_hi = min(c_hi, t_lo + per);
  unsigned long long tsum = 0;
  for (uint64_t v = t_lo; v < t_hi; ++v) tsum += C.deg[v];
  unsigned long long inc = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = shfl_up_u64(inc, d);
    if (lane >= static_cast<uint32_t>(d)) inc += o;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  unsigned long long wbase = 0;
#pragma unroll
  for (uint32_t j = 0; j < 8; ++j) wbase += (j < warp) ? warp_tot[j] : 0ull;
  const unsigned long long toff = wbase + inc - tsum;           // exclusive offset of this thread inside the CTA's chunk
  if (threadIdx.x == 255) C.cta_tot[blockIdx.x] = wbase + inc;
  grid.sync();
  {
    unsigned long long part = 0;
    for (uint32_t j = threadIdx.x; j < blockIdx.x; j += 256) part += C.cta_tot[j];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += shfl_xor_u64(part, d);
    __syncthreads();                                             // warp_tot is reused
    if (lane == 0) warp_tot[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = 0;
      for (int j = 0; j < 8; ++j) b += warp_tot[j];
      cta_base = b;
    }
    __syncthreads();
    unsigned long long run = cta_base + toff;
    for (uint64_t v = t_lo; v < t_hi; ++v) { C.row[v] = static_cast<uint32_t>(run); run += C.deg[v]; }
    if (tid == 0) C.row[n] = static_cast<uint32_t>(m);
  }
  grid.sync();
  cl_stamp(C, tid, slot);
  // ---- fill: links sorted by source (order inside a row is irrelevant)
  for (uint64_t e = tid; e < m; e += nth) {
    const uint2 ed = C.edges[e];
    const uint32_t pos = C.row[ed.x] + atomicSub(&C.deg[ed.x], 1u) - 1u;
    C.srcs[pos] = ed.x;
    C.dsts[pos] = ed.y;
  }
  grid.sync();
  cl_stamp(C, tid, slot);
  // ---- rounds: bits[r%3] = lowered in round r-1 (read), bits[(r+1)%3] = lowered in round r (set), bits[(r+2)%3] cleared
  uint32_t round = 0;
  for (;; ++round) {
    const uint32_t *rd = C.bits + static_cast<size_t>(round % 3) * C.nwords;
    uint32_t *wr = C.bits + static_cast<size_t>((round + 1) % 3) * C.nwords;
    uint32_t *cl = C.bits + static_cast<size_t>((round + 2) % 3) * C.nwords;
    if (tid == 0) C.flags[(round + 1) % 3] = 0;
    if (round) for (uint64_t w = tid; w < C.nwords; w += nth) cl[w] = 0;
    int ch = 0;
    // consecutive threads take consecutive links: sources ascend, so srcs / rd / key[src] reads coalesce
    for (uint64_t base = static_cast<uint64_t>(blockIdx.x) * blockDim.x * kCsU; base < m; base += nth * kCsU) {
      uint32_t s[kCsU], d[kCsU];
      bool act[kCsU];
#pragma unroll
      for (int k = 0; k < kCsU; ++k) {
        const uint64_t i = base + static_cast<uint64_t>(k) * blockDim.x + threadIdx.x;
        act[k] = i < m;
        s[k] = act[k] ? C.srcs[i] : 0u;
      }
      if (round) {
        uint32_t w[kCsU];
#pragma unroll
        for (int k = 0; k < kCsU; ++k) w[k] = act[k] ? rd[s[k] >> 5] : 0u;
#pragma unroll
        for (int k = 0; k < kCsU; ++k) act[k] = act[k] && ((w[k] >> (s[k] & 31u)) & 1u);
      }
      unsigned long long ks[kCsU], kd[kCsU];
#pragma unroll
      for (int k = 0; k < kCsU; ++k) {
        const uint64_t i = base + static_cast<uint64_t>(k) * blockDim.x + threadIdx.x;
        d[k] = act[k] ? C.dsts[i] : 0u;
        ks[k] = act[k] ? C.key[s[k]] : ~0ull;
      }
#pragma unroll
      for (int k = 0; k < kCsU; ++k) kd[k] = act[k] ? C.key[d[k]] : 0ull;
#pragma unroll
      for (int k = 0; k < kCsU; ++k) {
        const unsigned long long cand = ks[k] + 1ull;
        if (act[k] && cand < kd[k]) {
          if (atomicMin(&C.key[d[k]], cand) > cand) { atomicOr(&wr[d[k] >> 5], 1u << (d[k] & 31u)); ch = 1; }
        }
      }
    }
    if (__syncthreads_or(ch) && threadIdx.x == 0) C.flags[round % 3] = 1;
    grid.sync();
    cl_stamp(C, tid, slot);
    if (C.flags[round % 3] == 0) break;
  }
  // ---- parent = smallest predecessor one level up, then unpack
  for (uint64_t base = static_cast<uint64_t>(blockIdx.x) * blockDim.x * kCsU; base < m; base += nth * kCsU) {
    uint32_t s[kCsU], d[kCsU];
    unsigned long long ks[kCsU], kd[kCsU];
#pragma unroll
    for (int k = 0; k < kCsU; ++k) {
      const uint64_t i = base + static_cast<uint64_t>(k) * blockDim.x + threadIdx.x;
      s[k] = i < m ? C.srcs[i] : kNone;
      d[k] = i < m ? C.dsts[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kCsU; ++k) {
      ks[k] = s[k] != kNone ? C.key[s[k]] : 0ull;
      kd[k] = s[k] != kNone ? C.key[d[k]] : 0ull;
    }
#pragma unroll
    for (int k = 0; k < kCsU; ++k)
      if (s[k] != kNone && ks[k] + 1ull == kd[k]) atomicMin(&C.parent[d[k]], s[k]);
  }
  for (uint64_t v = tid; v < n; v += nth) {
    const unsigned long long kv = C.key[v];
    C.label[v] = static_cast<uint32_t>(kv >> 32);
    C.generation[v] = static_cast<uint32_t>(kv);
  }
  if (tid == 0 && C.rounds_out) *C.rounds_out = round + 1;
  if (C.ts) { grid.sync(); cl_stamp(C, tid, slot); if (tid == 0 && slot < 64) C.ts[slot] = 0; }
}

}  // namespace swb
