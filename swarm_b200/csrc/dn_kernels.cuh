// swarm_b200/csrc/dn_kernels.cuh — d>1 path: q-gram parity vectors, popcount prefilter, exact banded
// affine-cost global alignment with the reference's tie-breaks.
//
// Reference (/root/reference): db_qgrams_init -> findqgrams (src/db.cc:819-842, src/qgram.cc:68-96);
// qgram_diff = ceil(popcount(a^b)/10) (src/qgram.cc:247-252, src/popcnt.cc:45-62); search_do ->
// search8/search16 + backtrack<> (src/scan.cc:221-256, src/search16.cc:207-230, src/utils/backtrack.h)
// whose scalar statement is src/nw.cc:40-191; control loop src/algo.cc:384-602.
//
// The greedy loop is replaced by the same closed form as d=1: a directed link u->t
// exists iff (abundance(u) >= abundance(t) or -n) and the reference's optimal alignment of query u
// against target t has <= d differences; swarms/generation/parent then come from the d=1 clustering
// kernels.  The reference's prunings (q-gram bound, triangle inequality on `diffestimate`) are
// necessary conditions only; any valid necessary condition may be used to generate candidates, the
// accept decision is the exact DP below.
//
// Band: an alignment with <= d differences costs <= B = d*max(mismatch, gapopen+gapextend); any
// alignment that cheap has total gap length <= w = (B-gapopen)/gapextend, so it never leaves the
// diagonals [-w, +w].  Inside that band every cell on an optimal path, and every predecessor that ties
// with or beats the chosen one, is computed exactly (it lies on a path of cost <= B), so the direction
// flags along the backtrack equal the full-matrix ones.  If the banded score exceeds B the pair is
// rejected (the true optimum then exceeds B as well).
#pragma once
#include "d1_join.cuh"

namespace swb {

// out[0] = max, out[1] = ~min (both start at 0)
__global__ void k_minmax_u32(const uint32_t *v, uint32_t n, uint32_t *out) {
  uint32_t m = 0, w = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { m = max(m, v[i]); w = max(w, ~v[i]); }
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) { m = max(m, __shfl_xor_sync(kFull, m, k)); w = max(w, __shfl_xor_sync(kFull, w, k)); }
  if ((threadIdx.x & 31u) == 0) { atomicMax(out, m); atomicMax(out + 1, w); }
}

// *flag = 1 when some abundance is smaller than its successor's (the database is not sorted descending)
__global__ void k_unsorted_u64(const uint64_t *v, uint32_t n, uint32_t *flag) {
  bool bad = false;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x) bad |= v[i] < v[i + 1];
  if (__any_sync(kFull, bad) && (threadIdx.x & 31u) == 0) *flag = 1u;
}

// ---- q-gram parity vectors (a10) ------------------------------------------------------------------
// one warp per amplicon; the 1024-bit vector lives in shared memory (32 x u32), bits toggled with
// shared atomics, then written as one coalesced 128-byte row.
__global__ void __launch_bounds__(256) k_dn_qgrams(const uint64_t *words, const uint32_t *len, uint32_t stride, uint32_t n,
                                                   uint32_t *qgrams /* n * 32 */) {
  __shared__ uint32_t vec[8][32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < n; a += warps) {
    vec[warp][lane] = 0;
    __syncwarp();
    const uint64_t *w = words + static_cast<uint64_t>(a) * stride;
    const uint32_t L = len[a];
    for (uint32_t p = 4 + lane; p < L; p += 32) {
      uint32_t q = 0;
#pragma unroll
      for (uint32_t k = 0; k < 5; ++k) {
        const uint32_t pp = p - 4 + k;
        q = (q << 2) | (static_cast<uint32_t>(w[pp >> 5] >> ((pp & 31u) << 1)) & 3u);
      }
      atomicXor(&vec[warp][q >> 5], 1u << (q & 31u));       // bit index = qgram & 1023 (src/qgram.cc:91-93)
    }
    __syncwarp();
    qgrams[static_cast<uint64_t>(a) * 32 + lane] = vec[warp][lane];
    __syncwarp();
  }
}

struct DnParams {
  const uint64_t *words;
  const uint32_t *len;
  const uint64_t *abundance;
  const uint32_t *qgrams;
  uint32_t n, stride;
  uint32_t d;
  int ncb;
  int32_t mismatch, gapopen, gapextend;
  int32_t bound;       // B
  uint32_t w;          // band half-width
  uint32_t max_popc;   // 10*d
  uint2 *tasks;        // (query, target) alignment tasks
  unsigned long long *task_count;
  uint64_t task_cap;
  uint2 *edges;        // accepted links (src = query, dst = target)
  uint32_t *ediff;     // differences of each accepted link
  unsigned long long *edge_count;
  uint64_t edge_cap;
  uint32_t *dirs;      // scratch for direction flags: [row][word][thread]
  uint32_t dir_words;  // u32 words per row
  uint32_t max_len;
  unsigned long long *stats;   // [0] q-gram comparisons, [1] alignments, [2] pruned early
};

// ---- candidate generation (a10): tiled all-pairs popcount filter ------------------------------------
// Block = 256 targets (one per thread, vector in registers) x a tile of 32 queries in shared memory.
// Pair (u,t), u < t: task (query u, target t); if abundances tie (or -n) also (query t, target u)
// (src/algo.cc:423-431,518-531: the pool of a (sub)seed holds amplicons with abundance <= its own).
constexpr int kDnTileQ = 32;
__global__ void __launch_bounds__(256) k_dn_filter(DnParams P, uint32_t q_tile0, uint32_t n_qtiles) {
  __shared__ uint32_t qv[kDnTileQ][32];
  __shared__ uint32_t qlen_s[kDnTileQ];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool tin = t < P.n;
  uint32_t tv[32];
  if (tin) {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.qgrams + static_cast<uint64_t>(t) * 32);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint4 x = src[k]; tv[4 * k] = x.x; tv[4 * k + 1] = x.y; tv[4 * k + 2] = x.z; tv[4 * k + 3] = x.w; }
  }
  const uint32_t tlen = tin ? P.len[t] : 0u;
  const uint64_t tab = tin ? P.abundance[t] : 0ull;
  unsigned long long cmp = 0;
  const uint32_t t_block_last = min(P.n, (blockIdx.x + 1) * blockDim.x) - 1;
  for (uint32_t qt = q_tile0; qt < q_tile0 + n_qtiles; ++qt) {
    const uint32_t u0 = qt * kDnTileQ;
    if (u0 >= t_block_last) break;                         // only pairs u < t
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kDnTileQ * 32; i += blockDim.x) {
      const uint32_t u = u0 + (i >> 5);
      qv[i >> 5][i & 31] = u < P.n ? P.qgrams[static_cast<uint64_t>(u) * 32 + (i & 31)] : 0xFFFFFFFFu;
    }
    if (threadIdx.x < kDnTileQ) qlen_s[threadIdx.x] = (u0 + threadIdx.x) < P.n ? P.len[u0 + threadIdx.x] : 0u;
    __syncthreads();
    for (uint32_t j = 0; j < kDnTileQ; ++j) {
      const uint32_t u = u0 + j;
      bool cand = tin && u < t;
      if (cand) {
        const uint32_t ul = qlen_s[j];
        const uint32_t dl = ul > tlen ? ul - tlen : tlen - ul;
        if (dl > P.w) cand = false;
      }
      if (cand) {
        cmp++;
        uint32_t c = 0;
#pragma unroll
        for (int s = 0; s < 4 && c <= P.max_popc; ++s) {   // 256 bits per stage, stop as soon as hopeless
#pragma unroll
          for (int k = 0; k < 8; ++k) c += __popc(tv[8 * s + k] ^ qv[j][8 * s + k]);
        }
        cand = c <= P.max_popc;
      }
      // emit tasks (warp-aggregated append)
      const bool both = cand && (P.ncb || P.abundance[u] == tab);
      const uint32_t m1 = __ballot_sync(kFull, cand), m2 = __ballot_sync(kFull, both);
      if (m1) {
        const uint32_t lane = threadIdx.x & 31u;
        const uint32_t total = __popc(m1) + __popc(m2);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.task_count, static_cast<unsigned long long>(total));
        base = shfl_u64(base, 0);
        if (cand) {
          const unsigned long long at = base + __popc(m1 & ((1u << lane) - 1u));
          if (at < P.task_cap) P.tasks[at] = make_uint2(u, t);
        }
        if (both) {
          const unsigned long long at = base + __popc(m1) + __popc(m2 & ((1u << lane) - 1u));
          if (at < P.task_cap) P.tasks[at] = make_uint2(t, u);
        }
      }
    }
  }
  if (P.stats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) cmp += __shfl_xor_sync(kFull, cmp, m);
    if ((threadIdx.x & 31u) == 0 && cmp) atomicAdd(&P.stats[0], cmp);
  }
}

// ---- exact banded aligner (a11) -----------------------------------------------------------------------
// One (query, target) task per thread.  Recurrence, initial values and tie-breaks are those of
// src/nw.cc:40-112; direction flags use its polarity (1 up, 2 left, 4 ext-up, 8 ext-left) and the
// backtrack is src/nw.cc:115-191.  rows = target positions, columns = query positions.
constexpr int32_t kDnInf = 1 << 28;

template <int WMAX>
__global__ void __launch_bounds__(128) k_dn_align(DnParams P, uint64_t task0, uint64_t ntasks) {
  const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const int w = static_cast<int>(P.w);
  const int nb = 2 * w + 1;
  const int32_t go = P.gapopen, ge = P.gapextend, mis = P.mismatch;
  unsigned long long pruned = 0, done_cnt = 0;
  for (uint64_t task = task0 + gtid; task < task0 + ntasks; task += nthreads) {
    const uint2 tk = P.tasks[task];
    const uint32_t qid = tk.x, tid = tk.y;
    const int qlen = static_cast<int>(P.len[qid]), dlen = static_cast<int>(P.len[tid]);
    const uint64_t *qw = P.words + static_cast<uint64_t>(qid) * P.stride;
    const uint64_t *tw = P.words + static_cast<uint64_t>(tid) * P.stride;
    done_cnt++;
    // band state: index k <-> column c = r - w + k;  hb/eb hold the previous row's H/E at columns (r-1)-w+k
    int32_t hb[2 * WMAX + 2], eb[2 * WMAX + 2];
#pragma unroll
    for (int k = 0; k < 2 * WMAX + 2; ++k) {
      const int c = -1 - w + k;                          // "row -1": the DP boundary (src/nw.cc:64-68)
      const bool ok = k <= nb && c >= 0 && c < qlen;
      hb[k] = ok ? go + (c + 1) * ge : kDnInf;
      eb[k] = ok ? 2 * go + (c + 2) * ge : kDnInf;
    }
    // sliding window of query bases: bits 2k.. = base at column r - w + k
    uint64_t qwin = 0;
    for (int k = 0; k < nb; ++k) {
      const int c = -w + k;
      if (c >= 0 && c < qlen) qwin |= static_cast<uint64_t>(base_at(qw, static_cast<uint32_t>(c))) << (2 * k);
    }
    bool reject = false;
    uint64_t tword = 0;
    for (int r = 0; r < dlen; ++r) {
      if ((r & 31) == 0) tword = tw[r >> 5];
      const uint32_t tb = static_cast<uint32_t>(tword >> ((r & 31) << 1)) & 3u;
      const int c0 = r - w;
      int32_t top, diagonal;
      if (c0 <= 0) {                                       // true left boundary (src/nw.cc:75-76)
        top = 2 * go + (r + 2) * ge;
        diagonal = r == 0 ? 0 : go + r * ge;
      } else {
        top = kDnInf;                                      // cell (r, c0-1) is outside the band
        diagonal = hb[0];                                  // H(r-1, c0-1)
      }
      // columns c < 0 or c >= qlen do not exist: their band slots are skipped and never read
      // (a valid cell reads only its own column's previous-row slot k+1 and the carried diagonal)
      int32_t rowmin = kDnInf;
      uint32_t dw[(2 * WMAX + 1 + 7) / 8];
#pragma unroll
      for (int j = 0; j < (2 * WMAX + 1 + 7) / 8; ++j) dw[j] = 0;
#pragma unroll
      for (int k = 0; k < 2 * WMAX + 1; ++k) {
        const int c = c0 + k;
        if (k < nb && c >= 0 && c < qlen) {
          const int32_t prevdiag = hb[k + 1];              // H(r-1, c)
          int32_t left = eb[k + 1];                        // E(r-1, c)
          const uint32_t qb = static_cast<uint32_t>(qwin >> (2 * k)) & 3u;
          diagonal += (qb == tb) ? 0 : mis;
          uint32_t f = 0;
          if (top < diagonal) { f |= 1u; diagonal = top; }
          if (left < diagonal) diagonal = left;
          if (left == diagonal) f |= 2u;
          hb[k] = diagonal;
          rowmin = min(rowmin, diagonal);
          diagonal += go + ge;
          left += ge;
          top += ge;
          if (top < diagonal) f |= 4u;
          if (left < diagonal) f |= 8u;
          top = min(diagonal, top);
          left = min(diagonal, left);
          eb[k] = left;
          diagonal = prevdiag;
          dw[k >> 3] |= f << ((k & 7) * 4);
        }
      }
#pragma unroll
      for (int j = 0; j < (2 * WMAX + 1 + 7) / 8; ++j)
        if (j < static_cast<int>(P.dir_words)) P.dirs[(static_cast<uint64_t>(r) * P.dir_words + j) * nthreads + gtid] = dw[j];
      // slide the query window: new column r+1+w enters at index nb-1
      qwin >>= 2;
      {
        const int c = r + 1 + w;
        if (c < qlen) qwin |= static_cast<uint64_t>(base_at(qw, static_cast<uint32_t>(c))) << (2 * (nb - 1));
      }
#pragma unroll
      for (int k = 0; k < 2 * WMAX + 2; ++k)               // the column entering the band has no predecessor row
        if (k == nb) { hb[k] = kDnInf; eb[k] = kDnInf; }
      if (rowmin > P.bound) { reject = true; pruned++; break; }
    }
    if (reject) continue;
    // final score = H(dlen-1, qlen-1): band index of column qlen-1 in row dlen-1
    const int kend = (qlen - 1) - (dlen - 1) + w;
    int32_t score = kDnInf;
#pragma unroll
    for (int k = 0; k < 2 * WMAX + 1; ++k) if (k == kend) score = hb[k];
    if (score > P.bound) continue;
    // backtrack (src/nw.cc:133-187)
    int column = qlen, row = dlen;
    uint32_t alength = 0, matches = 0;
    int op = 0;                                            // 0 none, 1 'I', 2 'D', 3 'M'
    while (column > 0 && row > 0) {
      const int r = row - 1, c = column - 1;
      const int k = c - r + w;
      const uint32_t wv = P.dirs[(static_cast<uint64_t>(r) * P.dir_words + (k >> 3)) * nthreads + gtid];
      const uint32_t cell = (wv >> ((k & 7) * 4)) & 15u;
      ++alength;
      if (op == 1 && (cell & 8u)) { --row; }
      else if (op == 2 && (cell & 4u)) { --column; }
      else if (cell & 2u) { --row; op = 1; }
      else if (cell & 1u) { --column; op = 2; }
      else {
        if (base_at(qw, static_cast<uint32_t>(c)) == base_at(tw, static_cast<uint32_t>(r))) ++matches;
        --column; --row; op = 3;
      }
    }
    alength += static_cast<uint32_t>(column + row);
    const uint32_t diffs = alength - matches;
    if (diffs <= P.d) {
      const unsigned long long at = atomicAdd(P.edge_count, 1ull);
      if (at < P.edge_cap) { P.edges[at] = make_uint2(qid, tid); P.ediff[at] = diffs; }
    }
  }
  if (P.stats) {
    if (done_cnt) atomicAdd(&P.stats[1], done_cnt);
    if (pruned) atomicAdd(&P.stats[2], pruned);
  }
}

// ---- the same aligner for ANY band half-width (d > 6 with the default scoring: w > 15) ------------------------------
// k_dn_align keeps the band of a row in registers, which stops at w = 15.  Here the previous row's H and E live in a
// per-thread global scratch indexed by COLUMN (interleaved over threads, so a warp's accesses coalesce) and the four
// direction flags of cell (r, c) go to nibble c of row r: the recurrence, the boundary values, the band rule (cells
// outside |c - r| <= w are infinite, the band argument of DESIGN.md §3.6 makes that exact for every accepted pair) and
// the backtrack are the ones of k_dn_align; with w >= the sequence length this is the full matrix of src/nw.cc.
// Slower per cell (two loads and two stores), used only where the register kernel cannot go.
__global__ void __launch_bounds__(128) k_dn_align_wide(DnParams P, int32_t *hrow, int32_t *erow, uint64_t task0, uint64_t ntasks) {
  const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const int w = static_cast<int>(P.w);
  const int32_t go = P.gapopen, ge = P.gapextend, mis = P.mismatch;
  unsigned long long pruned = 0, done_cnt = 0;
  for (uint64_t task = task0 + gtid; task < task0 + ntasks; task += nthreads) {
    const uint2 tk = P.tasks[task];
    const uint32_t qid = tk.x, tid = tk.y;
    const int qlen = static_cast<int>(P.len[qid]), dlen = static_cast<int>(P.len[tid]);
    const uint64_t *qw = P.words + static_cast<uint64_t>(qid) * P.stride;
    const uint64_t *tw = P.words + static_cast<uint64_t>(tid) * P.stride;
    done_cnt++;
    for (int c = 0; c < qlen; ++c) {                       // "row -1": the DP boundary (src/nw.cc:64-68), inside its band
      const bool ok = c <= w;
      hrow[static_cast<uint64_t>(c) * nthreads + gtid] = ok ? go + (c + 1) * ge : kDnInf;
      erow[static_cast<uint64_t>(c) * nthreads + gtid] = ok ? 2 * go + (c + 2) * ge : kDnInf;
    }
    bool reject = false;
    for (int r = 0; r < dlen; ++r) {
      const uint32_t tb = base_at(tw, static_cast<uint32_t>(r));
      const int c_lo = max(0, r - w), c_hi = min(qlen - 1, r + w);
      int32_t top, diagonal;
      if (r - w <= 0) {                                    // true left boundary (src/nw.cc:75-76)
        top = 2 * go + (r + 2) * ge;
        diagonal = r == 0 ? 0 : go + r * ge;
      } else {
        top = kDnInf;                                      // cell (r, c_lo - 1) is outside the band
        diagonal = hrow[static_cast<uint64_t>(c_lo - 1) * nthreads + gtid];   // H(r-1, c_lo-1)
      }
      int32_t rowmin = kDnInf;
      uint32_t dw = 0;
      for (int c = c_lo; c <= c_hi; ++c) {
        const uint64_t at = static_cast<uint64_t>(c) * nthreads + gtid;
        const int32_t prevdiag = hrow[at];                 // H(r-1, c)
        int32_t left = erow[at];                           // E(r-1, c)
        diagonal += (base_at(qw, static_cast<uint32_t>(c)) == tb) ? 0 : mis;
        uint32_t f = 0;
        if (top < diagonal) { f |= 1u; diagonal = top; }
        if (left < diagonal) diagonal = left;
        if (left == diagonal) f |= 2u;
        hrow[at] = diagonal;
        rowmin = min(rowmin, diagonal);
        diagonal += go + ge;
        left += ge;
        top += ge;
        if (top < diagonal) f |= 4u;
        if (left < diagonal) f |= 8u;
        top = min(diagonal, top);
        left = min(diagonal, left);
        erow[at] = left;
        diagonal = prevdiag;
        dw |= f << ((c & 7) * 4);
        if ((c & 7) == 7 || c == c_hi) {
          P.dirs[(static_cast<uint64_t>(r) * P.dir_words + (c >> 3)) * nthreads + gtid] = dw;
          dw = 0;
        }
      }
      if (c_lo > c_hi || rowmin > P.bound) { reject = true; pruned++; break; }
    }
    if (reject) continue;
    if (abs(qlen - dlen) > w) continue;                    // the corner cell is outside the band
    if (hrow[static_cast<uint64_t>(qlen - 1) * nthreads + gtid] > P.bound) continue;
    // backtrack (src/nw.cc:133-187)
    int column = qlen, row = dlen;
    uint32_t alength = 0, matches = 0;
    int op = 0;                                            // 0 none, 1 'I', 2 'D', 3 'M'
    while (column > 0 && row > 0) {
      const int r = row - 1, c = column - 1;
      const uint32_t wv = P.dirs[(static_cast<uint64_t>(r) * P.dir_words + (c >> 3)) * nthreads + gtid];
      const uint32_t cell = (wv >> ((c & 7) * 4)) & 15u;
      ++alength;
      if (op == 1 && (cell & 8u)) { --row; }
      else if (op == 2 && (cell & 4u)) { --column; }
      else if (cell & 2u) { --row; op = 1; }
      else if (cell & 1u) { --column; op = 2; }
      else {
        if (base_at(qw, static_cast<uint32_t>(c)) == base_at(tw, static_cast<uint32_t>(r))) ++matches;
        --column; --row; op = 3;
      }
    }
    alength += static_cast<uint32_t>(column + row);
    const uint32_t diffs = alength - matches;
    if (diffs <= P.d) {
      const unsigned long long at = atomicAdd(P.edge_count, 1ull);
      if (at < P.edge_cap) { P.edges[at] = make_uint2(qid, tid); P.ediff[at] = diffs; }
    }
  }
  if (P.stats) {
    if (done_cnt) atomicAdd(&P.stats[1], done_cnt);
    if (pruned) atomicAdd(&P.stats[2], pruned);
  }
}

// ---- candidate generation by pigeonhole join (preferred; the all-pairs k_dn_filter is the fallback) --------
// <= d differences leave at least one of d+1 disjoint pieces of t untouched: P_j = t[jK,(j+1)K) for j < d and
// the suffix piece t[Lt-K, Lt), K = min(64, minlen/(d+1)).  Every amplicon stores its d+1 piece hashes; every
// amplicon a looks up: its prefix piece, its suffix piece, and for the middle pieces j = 1..d-1 the K-mers at
// offsets jK+δ, |δ| <= d.  A hit v > a (each unordered pair from its smaller id) that also passes the length
// and q-gram tests (src/qgram.cc:247-252) becomes alignment task (a -> v), plus (v -> a) when the abundances
// tie or with -n.  A pair reachable through several pieces is de-duplicated with a small hash set.
struct DnJoinParams {
  unsigned long long *table;
  uint64_t n_buckets;
  uint32_t K;
  unsigned long long *seen;      // open-addressing set of (a<<32|v)
  uint64_t seen_mask;
};

__global__ void __launch_bounds__(256) k_dn_index_pieces(DnParams P, DnJoinParams Q) {
  const uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t np = P.d + 1;
  const uint32_t a = static_cast<uint32_t>(t / np), piece = static_cast<uint32_t>(t % np);
  if (a >= P.n) return;
  const uint64_t *w = P.words + static_cast<uint64_t>(a) * P.stride;
  const uint32_t L = P.len[a];
  const uint32_t off = piece == P.d ? L - Q.K : piece * Q.K;
  const uint64_t h = piece_hash(w, P.stride, off, Q.K, piece);
  const unsigned long long val = (h << 32) | a;
  uint64_t b = __umul64hi(h, Q.n_buckets);
  for (;;) {
    unsigned long long *slot = Q.table + b * 4;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (slot[s] == kT2Empty && atomicCAS(&slot[s], kT2Empty, val) == kT2Empty) return;
    if (++b == Q.n_buckets) b = 0;
  }
}

__global__ void __launch_bounds__(256) k_dn_candidates_join(DnParams P, DnJoinParams Q) {
  __shared__ PairStage stage[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  PairStage &S = stage[warp];
  uint32_t scnt = 0;
  const uint32_t d = P.d, K = Q.K;
  const uint32_t nq = 2 + (d - 1) * (2 * d + 1);
  const uint64_t total = static_cast<uint64_t>(P.n) * nq;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t rounds = (total + nthreads - 1) / nthreads;
  unsigned long long cmp = 0;
  for (uint64_t r = 0; r < rounds; ++r) {
    const uint64_t t = r * nthreads + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t a = static_cast<uint32_t>(t / nq), q = static_cast<uint32_t>(t % nq);
    bool walking = t < total;
    uint32_t L = 0, tag = 0;
    uint64_t b = 0, aab = 0;
    if (walking) {
      const uint64_t *w = P.words + static_cast<uint64_t>(a) * P.stride;
      L = P.len[a];
      aab = P.abundance[a];
      uint32_t piece, off;
      if (q == 0) { piece = 0; off = 0; }
      else if (q == 1) { piece = d; off = L - K; }
      else {
        const uint32_t m = q - 2;
        piece = 1 + m / (2 * d + 1);
        const int o = static_cast<int>(piece * K) + static_cast<int>(m % (2 * d + 1)) - static_cast<int>(d);
        if (o < 0 || static_cast<uint32_t>(o) + K > L) walking = false;
        off = static_cast<uint32_t>(o < 0 ? 0 : o);
      }
      if (walking) {
        const uint64_t h = piece_hash(w, P.stride, off, K, piece);
        tag = static_cast<uint32_t>(h);
        b = __umul64hi(h, Q.n_buckets);
      }
    }
    while (__any_sync(kFull, walking)) {
      unsigned long long sv[4] = {kT2Empty, kT2Empty, kT2Empty, kT2Empty};
      if (walking) {
        const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(Q.table + b * 4);
        const ulonglong2 x = bp[0], y = bp[1];
        sv[0] = x.x; sv[1] = x.y; sv[2] = y.x; sv[3] = y.y;
      }
      bool full = walking;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t mine = 0;
        uint2 t1 = make_uint2(0, 0), t2 = make_uint2(0, 0);
        if (full) {
          if (sv[s] == kT2Empty) full = false;
          else if (static_cast<uint32_t>(sv[s] >> 32) == tag) {
            const uint32_t v = static_cast<uint32_t>(sv[s]);
            if (v > a) {
              const uint32_t Lv = P.len[v];
              const uint32_t dl = Lv > L ? Lv - L : L - Lv;
              bool cand = dl <= P.w;
              if (cand) {                                             // first time this pair is seen?
                const unsigned long long key = (static_cast<unsigned long long>(a) << 32) | v;
                uint64_t i = (key * 0x9E3779B97F4A7C15ull) >> 20 & Q.seen_mask;
                for (int probe = 0; probe < 64; ++probe) {
                  const unsigned long long old = atomicCAS(&Q.seen[i], kT2Empty, key);
                  if (old == kT2Empty) break;                         // inserted: new pair
                  if (old == key) { cand = false; break; }            // duplicate
                  i = (i + 1) & Q.seen_mask;
                }
              }
              if (cand) {                                             // q-gram lower bound
                cmp++;
                const uint4 *qa = reinterpret_cast<const uint4 *>(P.qgrams + static_cast<uint64_t>(a) * 32);
                const uint4 *qv = reinterpret_cast<const uint4 *>(P.qgrams + static_cast<uint64_t>(v) * 32);
                uint32_t c = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const uint4 x = qa[k], y = qv[k];
                  c += __popc(x.x ^ y.x) + __popc(x.y ^ y.y) + __popc(x.z ^ y.z) + __popc(x.w ^ y.w);
                }
                cand = c <= P.max_popc;
              }
              if (cand) {
                t1 = make_uint2(a, v); mine = 1;
                if (P.ncb || P.abundance[v] == aab) { t2 = make_uint2(v, a); mine = 2; }
              }
            }
          }
        }
        stage_push(S, scnt, mine, t1, t2, P.tasks, P.task_count, P.task_cap, lane);
      }
      if (walking) {
        if (!full) walking = false;
        else if (++b == Q.n_buckets) b = 0;
      }
    }
  }
  stage_flush(S, scnt, P.tasks, P.task_count, P.task_cap, lane);
  if (P.stats) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) cmp += __shfl_xor_sync(kFull, cmp, m);
    if (lane == 0 && cmp) atomicAdd(&P.stats[0], cmp);
  }
}

// differences between every amplicon and its parent (third column of the -i file, src/algo.cc:478,581)
__global__ void k_dn_pdiff(const uint2 *edges, const uint32_t *ediff, uint64_t m, const uint32_t *parent, uint32_t *pdiff) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < m; e += stride) {
    const uint2 ed = edges[e];
    if (parent[ed.y] == ed.x) pdiff[ed.y] = ediff[e];
  }
}

}  // namespace swb
