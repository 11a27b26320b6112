// swarm_b200/csrc/d1_tsroute.cuh — multi-GPU "Hashing sequences" for the tile store (BASELINE configs[4]:
// the table sharded by hash range): every rank hashes ONLY ITS OWN ROWS and ships each record to the rank that owns the
// record's tile, over NVLink peer memory, inside the kernel.
//   k_ts_route          a CTA stages 256 of the rank's packed rows in shared memory (one TMA bulk copy), hashes the two pieces
//                       of every row, counting-sorts the 512 records by owner rank in shared memory and appends every owner's
//                       run to THIS SENDER's sub-region of that owner's inbox: space is reserved with a LOCAL atomicAdd per
//                       owner and CTA (no remote atomics), the run is written with consecutive 8-byte stores (full NVLink
//                       sectors).  FAT records carry the packed row (entry + row: the database itself is sharded, no rank holds
//                       all rows); slim records are `entry, local tile` (16 bytes: the database is replicated and the join
//                       gathers rows locally).  The last CTA publishes the sender's counts and raises its "ready" flag in every
//                       peer's control block (system-scope release).
//   k_ts_wait           one warp: spins until every peer has raised a flag for this epoch (5 s timeout -> error, not a hang);
//                       stream order makes the following kernel wait with it, and a single waiting CTA cannot starve anybody.
//   k_ts_scatter_inbox  files the records that arrived into the local tile slots (d1_tilestore.cuh: ts_append); the tile of a FAT
//                       record is recomputed from the row it carries.  The last CTA tells every peer that this inbox may be
//                       overwritten by the next epoch.
// The inbox shares its memory with the clustering inboxes of d1_dist.cuh (the two phases never overlap).
#pragma once
#include "d1_dist.cuh"
#include "d1_tilestore.cuh"

namespace swb {

// control words of the index exchange live behind the clustering control block in every rank's peer-visible buffer
struct TsCtl {
  unsigned long long cnt[kDistMaxWorld];          // [sender] records routed into this rank's inbox in the current epoch
  unsigned long long ready[kDistMaxWorld];        // [sender] epoch whose records (and count) have all been written
  unsigned long long freed[kDistMaxWorld];        // [receiver] epoch that receiver has finished reading out of ITS inbox ...
                                                  // ... (written into every sender's control block)
  unsigned long long cl_ready[kDistMaxWorld];     // [peer] clustering call that peer is about to launch (k_dist_rendezvous)
};
constexpr size_t kTsCtlOffset = 2048;             // inside the kDistCtlBytes control area

struct TsRouteParams {
  TileStoreParams J;
  uint32_t rank, world;
  uint32_t tiles_per_rank;                        // owner(tile) = tile / tiles_per_rank
  unsigned char *peer[kDistMaxWorld];             // peer-visible buffer of every rank
  uint64_t inbox_cap;                             // records one sender may put into one inbox
  uint32_t rec_words;                             // words per inbox record: FAT 1 + stride, slim 2
  unsigned long long *counters;                   // [world] this sender's append counters (local memory)
  uint32_t *done_ctas;                            // last-CTA detection
  unsigned long long epoch;
  uint32_t *err;                                  // [0] timeout [1] inbox overflow
};

__device__ __forceinline__ TsCtl *ts_ctl(const TsRouteParams &R, uint32_t r) { return reinterpret_cast<TsCtl *>(R.peer[r] + kTsCtlOffset); }
__device__ __forceinline__ unsigned long long *ts_inbox(const TsRouteParams &R, uint32_t owner, uint32_t sender) {
  return reinterpret_cast<unsigned long long *>(R.peer[owner] + kDistCtlBytes) + static_cast<uint64_t>(sender) * R.inbox_cap * R.rec_words;
}

// which == 0: wait until every peer's inbox is free for `epoch` (freed >= epoch - 1); which == 1: until every sender's records
// of `epoch` have arrived (ready >= epoch)
__global__ void __launch_bounds__(32) k_ts_wait(TsRouteParams R, int which) {
  const uint32_t lane = threadIdx.x;
  if (lane >= R.world) return;
  volatile unsigned long long *w = which ? &ts_ctl(R, R.rank)->ready[lane] : &ts_ctl(R, R.rank)->freed[lane];
  const unsigned long long need = which ? R.epoch : R.epoch - 1;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  while (*w < need) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (t1 - t0 > 5000000000ull) { R.err[0] = 1u; break; }
  }
  __threadfence_system();
}

template <bool FAT>
__global__ void __launch_bounds__(kTsRows) k_ts_route(TsRouteParams R) {
  extern __shared__ __align__(128) unsigned char ts_smem[];
  const TileStoreParams &J = R.J;
  uint64_t *rows = reinterpret_cast<uint64_t *>(ts_smem);                                    // kTsRows * stride
  unsigned long long *stage = reinterpret_cast<unsigned long long *>(rows + static_cast<size_t>(kTsRows) * J.stride + 2);   // 2 * kTsRows records
  __shared__ uint64_t bar;
  __shared__ uint32_t cnt[kDistMaxWorld + 1], start[kDistMaxWorld + 1];
  __shared__ unsigned long long base[kDistMaxWorld];
  __shared__ uint32_t is_last;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t r0 = blockIdx.x * kTsRows;
  const uint32_t n_rows = min(kTsRows, J.row_count - r0);
  if (tid <= kDistMaxWorld) cnt[tid] = 0;
  ts_stage_rows(J, rows, &bar, r0, n_rows);              // includes a __syncthreads before the copy is issued

  const uint32_t rw = R.rec_words;
  unsigned long long e[2] = {0, 0};
  uint32_t owner[2] = {kNone, kNone}, tl[2] = {0, 0}, rk[2] = {0, 0};
  const uint64_t *w = rows + static_cast<size_t>(tid) * J.stride;
  uint32_t kept_local = 0;
  if (tid < n_rows) {
    const uint32_t r = r0 + tid;
    const uint32_t L = J.len[r];
    const uint64_t ab = J.abundance[r];
#pragma unroll
    for (uint32_t piece = 0; piece < 2; ++piece) {
      const uint64_t h = piece_hash(w, J.stride, piece ? L - J.K : 0u, J.K, piece);
      const uint32_t tile = static_cast<uint32_t>(__umul64hi(h, static_cast<uint64_t>(J.n_tiles)));
      owner[piece] = tile / R.tiles_per_rank;
      tl[piece] = tile - owner[piece] * R.tiles_per_rank;
      e[piece] = ts_pack(J, h, piece, L, ab, J.row_first + r);
      if (owner[piece] == R.rank) {                        // my own tile: no detour through the inbox
        const uint64_t ref = kTsRefLocal | (static_cast<uint64_t>(r) * J.stride);      // indirect tiles: the row stays where it is
        ts_append(J, tl[piece], e[piece], J.row_base ? &ref : w);
        owner[piece] = kNone;
        ++kept_local;
      }
    }
  }
  if (J.stats) {
    kept_local = __reduce_add_sync(kFull, kept_local);
    if (lane == 0 && kept_local) atomicAdd(J.appended, static_cast<unsigned long long>(kept_local));
  }
  // rank of every record inside its owner's run: one shared-memory atomic per warp and owner
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t peers = __match_any_sync(kFull, owner[k]);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t b = 0;
    if (lane == leader && owner[k] != kNone) b = atomicAdd(&cnt[owner[k]], static_cast<uint32_t>(__popc(peers)));
    b = __shfl_sync(kFull, b, leader);
    rk[k] = b + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t o = 0; o < R.world; ++o) { start[o] = run; run += cnt[o]; }
    start[R.world] = run;
  }
  if (tid < R.world && cnt[tid]) base[tid] = atomicAdd(&R.counters[tid], static_cast<unsigned long long>(cnt[tid]));
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (owner[k] != kNone) {
      unsigned long long *d = stage + static_cast<size_t>(start[owner[k]] + rk[k]) * rw;
      d[0] = e[k];
      if (FAT) { for (uint32_t x = 0; x < J.stride; ++x) d[1 + x] = w[x]; }
      else d[1] = tl[k];
    }
  __syncthreads();
  const uint32_t total_words = start[R.world] * rw;
  for (uint32_t u = tid; u < total_words; u += kTsRows) {
    const uint32_t rec = u / rw, x = u - rec * rw;
    uint32_t o = 0;
    while (rec >= start[o + 1]) ++o;
    const unsigned long long slot = base[o] + (rec - start[o]);
    if (slot >= R.inbox_cap) { R.err[1] = 1u; continue; }
    ts_inbox(R, o, R.rank)[slot * rw + x] = stage[u];
  }
  // the last CTA of this sender publishes its counts and raises `ready` everywhere
  __threadfence_system();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(R.done_ctas, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (is_last && tid < R.world) {
    __threadfence();
    TsCtl *pc = ts_ctl(R, tid);
    *reinterpret_cast<volatile unsigned long long *>(&pc->cnt[R.rank]) = *reinterpret_cast<volatile unsigned long long *>(&R.counters[tid]);
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(&pc->ready[R.rank]) = R.epoch;
    __threadfence_system();
    if (tid == 0) *R.done_ctas = 0;
  }
}

// grid-stride over the records of every sender's sub-region of MY inbox
template <bool FAT>
__global__ void __launch_bounds__(256) k_ts_scatter_inbox(TsRouteParams R) {
  const TileStoreParams &J = R.J;
  const uint32_t rw = R.rec_words;
  const TsCtl *me = ts_ctl(R, R.rank);
  const uint64_t nth = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned long long kept = 0;
  for (uint32_t s = 0; s < R.world; ++s) {
    if (s == R.rank) continue;                             // my own records went straight into the tiles
    const uint64_t n_rec = min(*reinterpret_cast<const volatile unsigned long long *>(&me->cnt[s]), static_cast<unsigned long long>(R.inbox_cap));
    const unsigned long long *in = ts_inbox(R, R.rank, s);
    for (uint64_t i = tid; i < n_rec; i += nth) {
      const unsigned long long *rec = in + i * rw;
      const unsigned long long e = __ldcg(rec);
      uint32_t t;
      if (FAT) {
        const uint32_t piece = ts_key(J, e) & 1u, L = ts_len(J, e);
        const uint64_t h = piece_hash(ts_u64(rec + 1), J.stride, piece ? L - J.K : 0u, J.K, piece);
        t = static_cast<uint32_t>(__umul64hi(h, static_cast<uint64_t>(J.n_tiles))) - J.t_lo;
      } else {
        t = static_cast<uint32_t>(__ldcg(rec + 1));
      }
      if (J.row_base) {                                   // indirect tiles: the record points at the row in this inbox
        const uint64_t ref = static_cast<uint64_t>((rec + 1) - J.row_base);
        ts_append(J, t, e, &ref);
      } else {
        ts_append(J, t, e, ts_u64(rec + 1));
      }
      ++kept;
    }
  }
  if (J.stats) {
    kept = __reduce_add_sync(kFull, static_cast<uint32_t>(kept));
    if ((threadIdx.x & 31u) == 0 && kept) atomicAdd(J.appended, kept);
  }
  // the last CTA tells every sender that this inbox has been read
  __shared__ uint32_t is_last;
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(R.done_ctas + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (is_last && threadIdx.x < R.world && !J.row_base) {      // (indirect tiles keep reading the inbox until the join is done: k_ts_release)
    *reinterpret_cast<volatile unsigned long long *>(&ts_ctl(R, threadIdx.x)->freed[R.rank]) = R.epoch;
    __threadfence_system();
    if (threadIdx.x == 0) R.done_ctas[1] = 0;
  }
  if (is_last && threadIdx.x == 0 && J.row_base) R.done_ctas[1] = 0;
}

// indirect tiles: the join has read the rows out of the inbox; tell every sender it may be overwritten by the next epoch
__global__ void __launch_bounds__(32) k_ts_release(TsRouteParams R) {
  if (threadIdx.x < R.world) {
    *reinterpret_cast<volatile unsigned long long *>(&ts_ctl(R, threadIdx.x)->freed[R.rank]) = R.epoch;
    __threadfence_system();
  }
}

// Ranks that SHARE one GPU (tests, swb200_dist_setup_local with repeated devices): the persistent clustering kernel of a rank
// fills its share of every SM and spins on the cross-rank barrier, so a peer that is still hashing or joining may not get an SM in
// a compatible shared-memory configuration — without the index exchange nothing else keeps the ranks in step.  One warp announces
// "about to cluster, call `epoch`" to every peer and waits for all of them; stream order holds the clustering kernel back.
__global__ void __launch_bounds__(32) k_dist_rendezvous(TsRouteParams R) {
  const uint32_t lane = threadIdx.x;
  if (lane >= R.world) return;
  *reinterpret_cast<volatile unsigned long long *>(&ts_ctl(R, lane)->cl_ready[R.rank]) = R.epoch;
  __threadfence_system();
  volatile unsigned long long *w = &ts_ctl(R, R.rank)->cl_ready[lane];
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  while (*w < R.epoch) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (t1 - t0 > 5000000000ull) { R.err[0] = 1u; break; }
  }
}

}  // namespace swb
