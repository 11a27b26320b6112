// swarm_b200/csrc/engine.cu — the C-ABI of include/swarm_b200.h: context, device memory, phase
// orchestration, CUDA-event timing.  Kernels live in d1_kernels.cuh (+ fastidious / d>1 files).
// There is NO CPU fallback in this library: every entry point runs CUDA kernels or fails.
#include "../../include/swarm_b200.h"
#include "d1_kernels.cuh"
#include "d1_network_v2.cuh"
#include "d1_fastidious.cuh"
#include "d1_fastidious_join.cuh"
#include "d1_join.cuh"
#include "d1_tilejoin.cuh"
#include "d1_tilestore.cuh"
#include "d1_frontier.cuh"
#include "d1_tsroute.cuh"
#include "d1_bucket.cuh"
#include "d1_cluster.cuh"
#include "d1_dist.cuh"
#include "dn_kernels.cuh"
#include "d0_derep.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

using namespace swb;

namespace {

thread_local std::string g_err;

struct CudaFail {
  int code;
};

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
      throw CudaFail{e_ == cudaErrorMemoryAllocation ? SWB200_ENOMEM : SWB200_ECUDA};              \
    }                                                                                              \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    CK(cudaMalloc(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

uint64_t splitmix64(uint64_t &x) {
  x += 0x9e3779b97f4a7c15ULL;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

// smallest power of two >= 10(n+1)/7 — same sizing rule as the reference
// (src/utils/hashtable_size.cc:29-42), computed in integers.
uint64_t table_slots(uint64_t n) {
  const uint64_t want = 10 * (n + 1) / 7;
  uint64_t s = 2;
  while (s < want) s <<= 1;
  return s;
}

}  // namespace

// expansion kernels of swb200_load_db_compact: 16-bit lengths -> u32, abundance runs -> one value per amplicon
__global__ void __launch_bounds__(256) k_expand_len16(const uint16_t *in, uint32_t *out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void __launch_bounds__(256) k_expand_runs(const uint64_t *value, const uint32_t *start, uint32_t n_runs, uint64_t *out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t lo = 0, hi = n_runs;                          // last run with start[run] <= i
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (start[mid] <= i) lo = mid; else hi = mid;
  }
  out[i] = value[lo];
}

struct swb200_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // options
  int enum_mode = SWB200_ENUM_JOIN;
  int bloom_bytes_per_slot = 1;
  int collect_stats = 0;
  int shard_rank = 0, shard_world = 1;
  int net_kernel = 0;   // 0 auto, 1 = k_d1_network (v1), 2 = k_d1_network_half (v2, HALF only)
  // database
  uint32_t n = 0, n_padded = 0, stride = 0, longest = 0, batch = 2, zlen = 0;
  DevBuf<uint64_t> words, abundance, ztab, hashes;
  DevBuf<uint32_t> len;
  std::vector<uint64_t> h_ztab;
  // index
  DevBuf<Slot> slots;
  DevBuf<uint2> filter;
  uint64_t n_slots = 0, n_filter_blocks = 0;
  bool indexed = false;
  // network
  DevBuf<uint2> edges;
  DevBuf<unsigned long long> counters;   // [0] edge_count, [1..4] stats, [8] dup flag(u32), [9] changed(u32)
  uint64_t n_edges = 0;
  bool have_network = false;
  int ncb = 0;
  // clustering
  DevBuf<uint32_t> label, generation, parent, cl_bits, cl_deg, cl_row, cl_srcs, cl_dsts;
  DevBuf<unsigned long long> cl_tot, cl_ts;
  // multi-GPU clustering over peer memory (d1_dist.cuh)
  uint32_t dist_rank = 0, dist_world = 0;
  unsigned char *dist_peer[kDistMaxWorld] = {};
  uint64_t dist_cap = 0;
  DevBuf<unsigned long long> dist_lcnt;
  DevBuf<uint2> dist_links;
  unsigned long long dist_calls = 0;
  uint32_t dist_rounds = 0;
  DevBuf<unsigned long long> key;
  bool clustered = false;
  // fastidious
  DevBuf<unsigned long long> mass, t2;
  DevBuf<uint32_t> light_ids, heavy_ids, graft;
  uint32_t max_len = 0, min_len = 0;
  uint32_t minmax[2] = {0, 0};
  uint32_t unsorted = 0;
  bool too_long = false;             // longest sequence > 5 000 nt: d = 0 only
  bool db_pending = false;           // rows uploaded by load_db_shard, exchange + db_commit still due
  bool sorted_desc = false;          // abundances never increase with the id (the reference's order, src/db.cc:392-406)
  int cluster_kernel = 0; // 0 (= 5) one persistent cooperative kernel that walks the link list every round (d1_kernels.cuh: k_cluster_persistent,
                          // the default: fastest on one GPU); 6 links bucketed by source block (d1_bucket.cuh, the multi-GPU kernel at world = 1);
                          // 4 frontier relaxation over out-rows (d1_frontier.cuh); 3 links first sorted by source (d1_cluster.cuh);
                          // 2 one launch per round; 1 label propagation + BFS
  int cluster_pack = 1;  // k_cluster_persistent / k_cluster_bucket relax ONE packed word swarm | generation | parent (no parent pass); 0 = r1's key + parent pass
  uint32_t cluster_gen_bits = 0;     // test hook: pretend the packed word has only this many generation bits (exercises the unpacked fallback)
  unsigned long long cluster_unpacked_reruns = 0;
  int cluster_hints = 1; // streaming cache policy for the link list / outputs of k_cluster_persistent (0 = plain loads, for comparison)
  int dn_filter = 0;     // 0 auto (pigeonhole join when possible), 1 = all-pairs q-gram filter
  int fast_kernel = 0;   // 0 auto, 1 = microvariant multimap (d1_fastidious.cuh), 2 = pigeonhole join
  uint32_t fast_chunks = 8;          // heavy amplicons are processed in this many ascending id chunks (pruning by graft_cand[l] <= h): 4.06 ms at 8, 4.29 at 16, 4.82 at 32 (10 M, r2z)
  DevBuf<uint8_t> is_light;
  DevBuf<unsigned long long> fj_bloom;
  DevBuf<uint2> cands;
  DevBuf<unsigned long long> jtab;   // K-mer multimap of the JOIN network
  uint64_t jtab_buckets = 0, jb_lo = 0, jb_hi = 0;
  uint32_t jK = 0;
  bool join_active = false;
  // partitioned (tile) join: d1_tilejoin.cuh
  int join_kernel = 0;               // 0 auto (tile store when a tile fits shared memory, d1_tilestore.cuh), 1 = global hash multimap
                                     // (d1_join.cuh), 2 = r1's count / scan / scatter tile join (d1_tilejoin.cuh)
  bool tile_active = false;
  // tile store (d1_tilestore.cuh): fat records in fixed-capacity tile slots
  bool ts_active = false;
  uint32_t ts_qcap = 1536, ts_outcap = 512;
  int ts_occ = 4, ts_occ_opt = 0;    // CTAs of the join per SM (register bound of the kernel variant): 4, 5 or 6
  uint32_t ts_cap_opt = 0;           // tuning: records per tile slot (0 = derived from the record size)
  int ts_fat = 0;                    // 1: records carry their packed row (the sharded-database multi-GPU layout); 0: 8-byte entries, rows gathered
  DevBuf<unsigned long long> ts_store, ts_ovf;
  DevBuf<uint32_t> ts_cursor;        // [T] cursors, then [T] overflow chain heads
  uint32_t ts_cap = 0, ts_tiles = 0, ts_lo = 0, ts_hi = 0;
  uint64_t ts_ovf_cap = 0;
  size_t ts_smem = 0;
  unsigned long long ts_overflow = 0, ts_fallbacks = 0;
  // multi-GPU index exchange (d1_tsroute.cuh) and sharded database
  bool db_sharded = false;           // this context holds only the rows [row_first, row_first + row_count) of the job's database
  uint32_t row_first = 0, row_count = 0;
  DevBuf<uint32_t> run_start_d;      // abundance runs of the WHOLE database (equal-abundance test without the abundance array)
  uint32_t n_runs = 0;
  uint32_t job_min_len = 0, job_max_len = 0;   // length range of the whole job (every rank must use the same piece length K)
  int dist_grid_div = 1;             // test hook: several ranks share ONE GPU, each persistent kernel takes 1/div of the SMs
  int index_exchange = 1;            // 0 never, 1 auto, 2 always:            // after swb200_dist_setup: hash only this rank's rows, route the records to the tile owners
  unsigned long long idx_epoch = 0;
  DevBuf<uint8_t> bk_flag;                     // bucketed clustering (d1_bucket.cuh)
  DevBuf<uint32_t> bk_count;
  DevBuf<unsigned long long> bk_off;
  DevBuf<uint2> bk_links, bk_unit;
  DevBuf<uint32_t> bk_act;
  int dist_kernel = 0;                         // 0 = links bucketed by source block (d1_bucket.cuh), 1 = r1's k_cluster_dist
  DevBuf<unsigned char> dist_own;              // peer-visible buffer allocated by swb200_dist_setup_local
  DevBuf<unsigned long long> ts_route_cnt;     // [16] sender counters, then done_ctas[2] + err[2] as 32-bit words
  uint64_t dist_buffer_bytes = 0;
  int skew_fallback = 1;             // dense data: abandon the quadratic overflow sweep for the linear enumeration (single GPU)
  // frontier clustering (d1_frontier.cuh)
  DevBuf<uint32_t> fr_deg, fr_adj;
  DevBuf<uint2> fr_spill;
  DevBuf<uint32_t> tj_count, tj_cursor, tj_big;
  DevBuf<unsigned long long> tj_off, tj_entries;
  DevBuf<uint2> tj_plist_ent;
  DevBuf<uint32_t> tj_plist_tile;
  DevBuf<unsigned long long> tj_plist_cnt;
  uint32_t tj_tiles = 0, tj_lo = 0, tj_hi = 0, tj_cmax = 0;
  uint32_t tj_cmax_override = 0;     // test hook: pretend shared memory holds fewer entries (exercises k_tile_join_big)
  unsigned long long tj_total = 0;
  size_t tj_smem = 0;
  uint64_t fstats[4] = {0, 0, 0, 0};
  // d>1
  DevBuf<uint32_t> qgrams, ediff, dirs, pdiff;
  DevBuf<uint2> tasks;
  uint64_t dnstats[4] = {0, 0, 0, 0};
  // compact loader staging
  DevBuf<uint16_t> ld_len16;
  DevBuf<uint64_t> ld_run_value;
  DevBuf<uint32_t> ld_run_start;
  // d = 0
  DevBuf<unsigned long long> dr_table, dr_mass;
  DevBuf<uint32_t> dr_slot, dr_rep, dr_size, dr_single;
  uint64_t drstats[3] = {0, 0, 0};
  // pinned staging
  void *pinned = nullptr;
  size_t pinned_bytes = 0;
  // timing
  double last_s = 0;
  double phase_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t launches = 0;
  uint64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  // multi-GPU "Hashing sequences": hash only this rank's rows and route the records to the tile owners (d1_tsroute.cuh)?  A sharded
  // database has no alternative; a replicated one can also be scanned whole by every rank, which is cheaper at 2 GPUs (2 x 0.34 ms
  // against 0.9 ms for route + inbox scatter, profiles/r2i) and loses from 3 GPUs on
  bool exchange() const { return dist_world > 1 && (db_sharded || index_exchange == 2 || (index_exchange == 1 && dist_world >= 3)); }
  void *staging(size_t bytes) {
    if (bytes > pinned_bytes) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      CK(cudaMallocHost(&pinned, bytes));
      pinned_bytes = bytes;
    }
    return pinned;
  }
  void tic() { CK(cudaEventRecord(ev0, stream)); }
  double toc(int phase) {
    CK(cudaEventRecord(ev1, stream));
    CK(cudaEventSynchronize(ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    last_s = ms * 1e-3;
    phase_s[phase] = last_s;
    return last_s;
  }
  D1Params params() {
    D1Params P{};
    P.words = words.p; P.len = len.p; P.abundance = abundance.p;
    P.n = n; P.stride = stride; P.batch = batch;
    P.ztab = ztab.p; P.zlen = zlen;
    P.slots = slots.p; P.slot_mask = n_slots - 1;
    P.filter = filter.p; P.filter_mask = static_cast<uint32_t>(n_filter_blocks - 1);
    P.hashes = hashes.p;
    P.edges = edges.p; P.edge_count = counters.p; P.edge_cap = edges.n;
    P.stats = collect_stats ? counters.p + 1 : nullptr;
    P.dup_flag = reinterpret_cast<uint32_t *>(counters.p + 8);
    P.no_cluster_breaking = ncb;
    if (const char *d = std::getenv("SWB200_DEBUG")) P.dbg = static_cast<uint32_t>(std::atoi(d));
    return P;
  }
  size_t network_smem() const {
    const size_t zwords = (static_cast<size_t>(zlen) * 4 + 1) & ~static_cast<size_t>(1);
    return zwords * 8 + static_cast<size_t>(kWarpsPerCta) * 2 * batch * stride * 8 +
           static_cast<size_t>(kWarpsPerCta) * sizeof(WarpScratch) + kWarpsPerCta * 2 * 8;
  }
};

#define API_BEGIN(ctx)                                       \
  if (!(ctx)) { g_err = "null context"; return SWB200_EINVAL; } \
  try {                                                      \
    CK(cudaSetDevice((ctx)->device));
#define API_END()                    \
  }                                  \
  catch (const CudaFail &f) {        \
    return f.code;                   \
  }                                  \
  catch (const std::bad_alloc &) {   \
    g_err = "host allocation failed"; \
    return SWB200_ENOMEM;            \
  }                                  \
  return SWB200_OK;

extern "C" {

const char *swb200_last_error(void) { return g_err.c_str(); }

int swb200_create(swb200_ctx **out, int device) {
  if (!out) { g_err = "null out pointer"; return SWB200_EINVAL; }
  *out = nullptr;
  swb200_ctx *c = nullptr;
  try {
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { g_err = "no such CUDA device"; return SWB200_EINVAL; }
    CK(cudaSetDevice(device));
    c = new swb200_ctx();
    c->device = device;
    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    c->counters.alloc(64);
  } catch (const CudaFail &f) {
    delete c;
    return f.code;
  }
  *out = c;
  return SWB200_OK;
}

void swb200_destroy(swb200_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  c->words.release(); c->abundance.release(); c->ztab.release(); c->hashes.release(); c->len.release();
  c->slots.release(); c->filter.release(); c->edges.release(); c->counters.release();
  c->label.release(); c->generation.release(); c->parent.release(); c->key.release(); c->cl_bits.release();
  c->cl_deg.release(); c->cl_row.release(); c->cl_srcs.release(); c->cl_dsts.release(); c->cl_tot.release(); c->cl_ts.release(); c->dist_lcnt.release(); c->dist_links.release();
  c->mass.release(); c->t2.release(); c->light_ids.release(); c->heavy_ids.release(); c->graft.release();
  c->is_light.release(); c->fj_bloom.release(); c->cands.release(); c->jtab.release();
  c->tj_count.release(); c->tj_cursor.release(); c->tj_big.release(); c->tj_off.release(); c->tj_entries.release(); c->tj_plist_ent.release(); c->tj_plist_tile.release(); c->tj_plist_cnt.release();
  c->run_start_d.release(); c->ts_route_cnt.release(); c->dist_own.release();
  c->bk_flag.release(); c->bk_count.release(); c->bk_off.release(); c->bk_links.release(); c->bk_unit.release(); c->bk_act.release();
  c->ts_store.release(); c->ts_ovf.release(); c->ts_cursor.release(); c->fr_deg.release(); c->fr_adj.release(); c->fr_spill.release();
  c->ld_len16.release(); c->ld_run_value.release(); c->ld_run_start.release();
  c->dr_table.release(); c->dr_mass.release(); c->dr_slot.release(); c->dr_rep.release(); c->dr_size.release(); c->dr_single.release();
  c->qgrams.release(); c->ediff.release(); c->dirs.release(); c->pdiff.release(); c->tasks.release();
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int swb200_set_option(swb200_ctx *c, const char *key, int64_t v) {
  if (!c || !key) { g_err = "null argument"; return SWB200_EINVAL; }
  const std::string k(key);
  if (k == "enum_mode" && (v == SWB200_ENUM_FULL || v == SWB200_ENUM_HALF || v == SWB200_ENUM_JOIN)) c->enum_mode = static_cast<int>(v);
  else if (k == "bloom_bytes_per_slot" && (v == 1 || v == 2 || v == 4 || v == 8)) c->bloom_bytes_per_slot = static_cast<int>(v);
  else if (k == "collect_stats") c->collect_stats = v != 0;
  else if (k == "net_kernel" && v >= 0 && v <= 2) c->net_kernel = static_cast<int>(v);
  else if (k == "join_kernel" && v >= 0 && v <= 2) c->join_kernel = static_cast<int>(v);
  else if (k == "skew_fallback" && (v == 0 || v == 1)) c->skew_fallback = static_cast<int>(v);
  else if (k == "tile_rows" && (v == 0 || v == 1)) c->ts_fat = static_cast<int>(v);
  else if (k == "index_exchange" && v >= 0 && v <= 2) c->index_exchange = static_cast<int>(v);
  else if (k == "dist_grid_div" && v >= 1 && v <= 16) c->dist_grid_div = static_cast<int>(v);
  else if (k == "dist_kernel" && (v == 0 || v == 1)) c->dist_kernel = static_cast<int>(v);
  else if (k == "job_min_len" && v >= 0 && v < (1ll << 32)) c->job_min_len = static_cast<uint32_t>(v);
  else if (k == "job_max_len" && v >= 0 && v < (1ll << 32)) c->job_max_len = static_cast<uint32_t>(v);
  else if (k == "tile_cap" && v >= 0 && v <= 768) c->ts_cap_opt = static_cast<uint32_t>(v);
  else if (k == "join_occupancy" && (v == 0 || (v >= 4 && v <= 6))) c->ts_occ_opt = static_cast<int>(v);
  else if (k == "tile_cmax" && v >= 0 && v <= 1024) c->tj_cmax_override = static_cast<uint32_t>(v);
  else if (k == "fast_kernel" && v >= 0 && v <= 2) c->fast_kernel = static_cast<int>(v);
  else if (k == "fast_chunks" && v >= 1 && v <= 1024) c->fast_chunks = static_cast<uint32_t>(v);
  else if (k == "dn_filter" && v >= 0 && v <= 1) c->dn_filter = static_cast<int>(v);
  else if (k == "cluster_kernel" && v >= 0 && v <= 6) c->cluster_kernel = static_cast<int>(v);
  else if (k == "cluster_hints" && (v == 0 || v == 1)) c->cluster_hints = static_cast<int>(v);
  else if (k == "cluster_pack" && (v == 0 || v == 1)) c->cluster_pack = static_cast<int>(v);
  else if (k == "cluster_gen_bits" && v >= 0 && v <= 32) c->cluster_gen_bits = static_cast<uint32_t>(v);
  else if (k == "shard_rank" && v >= 0) c->shard_rank = static_cast<int>(v);
  else if (k == "shard_world" && v >= 1) c->shard_world = static_cast<int>(v);
  else { g_err = "unknown option or bad value: " + k; return SWB200_EINVAL; }
  return SWB200_OK;
}

// device buffers for a database of n amplicons with `stride_words` words per row
static void db_alloc(swb200_ctx *c, uint32_t n, uint32_t stride_words) {
  c->indexed = c->have_network = c->clustered = false;
  c->db_sharded = false;
  c->row_first = 0;
  c->row_count = n;
  // seeds per TMA batch: even, <= kMaxBatch, batch*stride*8 bytes <= 2 KB per buffer
  uint32_t batch = 256 / stride_words;
  batch = std::max<uint32_t>(2, std::min<uint32_t>(kMaxBatch, batch & ~1u));
  c->batch = batch;
  c->n = n;
  c->stride = stride_words;
  c->n_padded = (n + batch - 1) / batch * batch;
  // room for an in-place all-gather of equal shards of ceil(n / world) rows (swb200_load_db_shard)
  const size_t shard = (static_cast<size_t>(n) + c->shard_world - 1) / c->shard_world;
  const size_t rows = std::max<size_t>(c->n_padded, shard * c->shard_world);
  c->words.alloc(rows * stride_words + 2);            // + 16 bytes: the TMA row staging rounds its size up to 16 (d1_tilestore.cuh)
  c->len.alloc(rows);
  c->abundance.alloc(rows);
}

// after the rows are on the device: padding, length range, order check, Zobrist table
static int db_finalize(swb200_ctx *c) {
  const uint32_t n = c->db_sharded ? c->row_count : c->n, stride_words = c->stride;      // rows present in this context
  if (!c->db_sharded && c->n_padded > n)
    CK(cudaMemsetAsync(c->words.p + static_cast<size_t>(n) * stride_words, 0,
                       static_cast<size_t>(c->n_padded - n) * stride_words * 8, c->stream));
  // Zobrist table: zlen = longest + 2 positions (two insertions, src/db.cc:652), 4 values each,
  // position-major.  Values come from splitmix64 — results do not depend on the hash function
  // (every hit is verified exactly).
  c->longest = stride_words * 32;
  c->zlen = c->longest + 2;
  // the clustering kernels keep per-position tables on chip; dereplication (d = 0) has no such limit
  c->too_long = static_cast<size_t>(c->zlen) * 32 > 160 * 1024;
  {
    CK(cudaMemsetAsync(c->counters.p + 15, 0, 8, c->stream));
    k_minmax_u32<<<c->sm_count * 2, 256, 0, c->stream>>>(c->len.p, n, reinterpret_cast<uint32_t *>(c->counters.p + 15));
    CK(cudaMemcpyAsync(c->minmax, c->counters.p + 15, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemsetAsync(c->counters.p + 20, 0, 8, c->stream));
    k_unsorted_u64<<<c->sm_count * 2, 256, 0, c->stream>>>(c->abundance.p, n, reinterpret_cast<uint32_t *>(c->counters.p + 20));
    CK(cudaMemcpyAsync(&c->unsorted, c->counters.p + 20, 4, cudaMemcpyDeviceToHost, c->stream));
    c->launches += 2;
  }
  if (!c->too_long && c->h_ztab.size() != static_cast<size_t>(c->zlen) * 4) {
    c->h_ztab.resize(static_cast<size_t>(c->zlen) * 4);
    uint64_t sm = 0x5eedb200c0ffeeULL;
    for (auto &z : c->h_ztab) z = splitmix64(sm);
    c->ztab.alloc(c->h_ztab.size());
    CK(cudaMemcpyAsync(c->ztab.p, c->h_ztab.data(), c->h_ztab.size() * 8, cudaMemcpyHostToDevice, c->stream));
  }
  c->toc(0);
  c->max_len = c->minmax[0];
  c->min_len = ~c->minmax[1];
  if (c->job_max_len) c->max_len = std::max(c->max_len, c->job_max_len);      // multi-GPU: the job's range, identical on every rank
  if (c->job_min_len) c->min_len = std::min(c->min_len, c->job_min_len);
  c->sorted_desc = c->unsorted == 0;
  return SWB200_OK;
}

// host -> device copy of `bytes`; pinned sources go straight to the copy engine, pageable ones through a pinned
// double buffer in 64 MiB chunks so the copies still run at the full PCIe rate
static void db_upload(swb200_ctx *c, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return;
  cudaPointerAttributes attr{};
  const bool is_pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (is_pinned) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return;
  }
  const size_t chunk = 64u << 20;
  char *stage = static_cast<char *>(c->staging(2 * chunk));
  size_t off = 0;
  int which = 0;
  cudaEvent_t done[2];
  CK(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
  bool used[2] = {false, false};
  while (off < bytes) {
    const size_t nb = std::min(chunk, bytes - off);
    if (used[which]) CK(cudaEventSynchronize(done[which]));
    std::memcpy(stage + which * chunk, static_cast<const char *>(src) + off, nb);
    CK(cudaMemcpyAsync(static_cast<char *>(dst) + off, stage + which * chunk, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(done[which], c->stream));
    used[which] = true;
    which ^= 1;
    off += nb;
  }
  CK(cudaStreamSynchronize(c->stream));
  cudaEventDestroy(done[0]);
  cudaEventDestroy(done[1]);
}

int swb200_load_db(swb200_ctx *c, const uint64_t *words, uint32_t stride_words, const uint32_t *len,
                   const uint64_t *abundance, uint32_t n) {
  API_BEGIN(c)
  if (!words || !len || !abundance || n == 0 || stride_words == 0) { g_err = "load_db: bad argument"; return SWB200_EINVAL; }
  if (n >= 0xFFFFFFF0u) { g_err = "load_db: too many amplicons"; return SWB200_EINVAL; }
  db_alloc(c, n, stride_words);
  c->tic();
  db_upload(c, c->words.p, words, static_cast<size_t>(n) * stride_words * 8);
  db_upload(c, c->len.p, len, static_cast<size_t>(n) * 4);
  db_upload(c, c->abundance.p, abundance, static_cast<size_t>(n) * 8);
  return db_finalize(c);
  API_END()
}

// same database as swb200_load_db from fewer host bytes: lengths as u16, abundances as runs (the database is sorted by
// abundance, so a run is one distinct value): 10 M x 150 bp = 420 MB over PCIe instead of 520 MB
int swb200_load_db_compact(swb200_ctx *c, const uint64_t *words, uint32_t stride_words, const uint16_t *len16,
                           const uint64_t *run_abundance, const uint32_t *run_start, uint32_t n_runs, uint32_t n) {
  API_BEGIN(c)
  if (!words || !len16 || !run_abundance || !run_start || n == 0 || n_runs == 0 || n_runs > n || stride_words == 0) {
    g_err = "load_db_compact: bad argument";
    return SWB200_EINVAL;
  }
  if (n >= 0xFFFFFFF0u) { g_err = "load_db_compact: too many amplicons"; return SWB200_EINVAL; }
  if (run_start[0] != 0 || run_start[n_runs] != n) { g_err = "load_db_compact: runs must cover [0, n)"; return SWB200_EINVAL; }
  for (uint32_t r = 0; r < n_runs; ++r)
    if (run_start[r] >= run_start[r + 1]) { g_err = "load_db_compact: empty or unordered run"; return SWB200_EINVAL; }
  db_alloc(c, n, stride_words);
  c->ld_len16.alloc(n); c->ld_run_value.alloc(n_runs); c->ld_run_start.alloc(static_cast<size_t>(n_runs) + 1);
  c->tic();
  db_upload(c, c->ld_len16.p, len16, static_cast<size_t>(n) * 2);
  db_upload(c, c->ld_run_value.p, run_abundance, static_cast<size_t>(n_runs) * 8);
  db_upload(c, c->ld_run_start.p, run_start, (static_cast<size_t>(n_runs) + 1) * 4);
  const unsigned blocks = (n + 255) / 256;
  k_expand_len16<<<blocks, 256, 0, c->stream>>>(c->ld_len16.p, c->len.p, n);
  k_expand_runs<<<blocks, 256, 0, c->stream>>>(c->ld_run_value.p, c->ld_run_start.p, n_runs, c->abundance.p, n);
  c->launches += 2;
  db_upload(c, c->words.p, words, static_cast<size_t>(n) * stride_words * 8);
  return db_finalize(c);
  API_END()
}

int swb200_load_db_shard(swb200_ctx *c, const uint64_t *words, uint32_t stride_words, const uint32_t *len,
                         const uint64_t *abundance, uint32_t n_total, uint32_t first, uint32_t count) {
  API_BEGIN(c)
  if (n_total == 0 || stride_words == 0 || static_cast<uint64_t>(first) + count > n_total || (count && (!words || !len || !abundance))) {
    g_err = "load_db_shard: bad argument";
    return SWB200_EINVAL;
  }
  if (n_total >= 0xFFFFFFF0u) { g_err = "load_db_shard: too many amplicons"; return SWB200_EINVAL; }
  {                                     // the in-place all-gather writes shard_world equal shards of ceil(n / world) rows: the shard must be one of them
    const uint64_t per = (static_cast<uint64_t>(n_total) + c->shard_world - 1) / c->shard_world;
    if (first % per != 0 || count > per) { g_err = "load_db_shard: [first, first + count) is not a shard of ceil(n / shard_world) rows (set \"shard_world\" first)"; return SWB200_EINVAL; }
  }
  db_alloc(c, n_total, stride_words);
  c->tic();
  db_upload(c, c->words.p + static_cast<size_t>(first) * stride_words, words, static_cast<size_t>(count) * stride_words * 8);
  db_upload(c, c->len.p + first, len, static_cast<size_t>(count) * 4);
  db_upload(c, c->abundance.p + first, abundance, static_cast<size_t>(count) * 8);
  c->db_pending = true;
  API_END()
}

// swb200_load_db_shard from fewer host bytes: 16-bit lengths for the rank's rows, and NO abundance upload at all — the abundance
// runs of the whole database (a few KB) are expanded on every device, so the row exchange only has to move words and lengths
int swb200_load_db_shard_compact(swb200_ctx *c, const uint64_t *words, uint32_t stride_words, const uint16_t *len16, uint32_t n_total,
                                 uint32_t first, uint32_t count, const uint64_t *run_abundance, const uint32_t *run_start, uint32_t n_runs) {
  API_BEGIN(c)
  if (n_total == 0 || stride_words == 0 || static_cast<uint64_t>(first) + count > n_total || (count && (!words || !len16)) ||
      !run_abundance || !run_start || n_runs == 0 || n_runs > n_total || run_start[0] != 0 || run_start[n_runs] != n_total || n_total >= 0xFFFFFFF0u) {
    g_err = "load_db_shard_compact: bad argument";
    return SWB200_EINVAL;
  }
  for (uint32_t r = 0; r < n_runs; ++r)
    if (run_start[r] >= run_start[r + 1]) { g_err = "load_db_shard_compact: empty or unordered run"; return SWB200_EINVAL; }
  {
    const uint64_t per = (static_cast<uint64_t>(n_total) + c->shard_world - 1) / c->shard_world;
    if (first % per != 0 || count > per) { g_err = "load_db_shard_compact: [first, first + count) is not a shard of ceil(n / shard_world) rows (set \"shard_world\" first)"; return SWB200_EINVAL; }
  }
  db_alloc(c, n_total, stride_words);
  c->ld_len16.alloc(std::max<uint32_t>(count, 1)); c->ld_run_value.alloc(n_runs); c->ld_run_start.alloc(static_cast<size_t>(n_runs) + 1);
  c->tic();
  db_upload(c, c->ld_len16.p, len16, static_cast<size_t>(count) * 2);
  db_upload(c, c->ld_run_value.p, run_abundance, static_cast<size_t>(n_runs) * 8);
  db_upload(c, c->ld_run_start.p, run_start, (static_cast<size_t>(n_runs) + 1) * 4);
  if (count) k_expand_len16<<<(count + 255) / 256, 256, 0, c->stream>>>(c->ld_len16.p, c->len.p + first, count);
  k_expand_runs<<<(n_total + 255) / 256, 256, 0, c->stream>>>(c->ld_run_value.p, c->ld_run_start.p, n_runs, c->abundance.p, n_total);
  c->launches += 2;
  db_upload(c, c->words.p + static_cast<size_t>(first) * stride_words, words, static_cast<size_t>(count) * stride_words * 8);
  CK(cudaGetLastError());
  c->db_pending = true;
  API_END()
}

// Sharded database (BASELINE configs[4]): this context keeps ONLY the rows [first, first + count) of the job's sorted database;
// the abundance runs of the whole database (tiny: one entry per distinct abundance) are replicated so that any rank can
// tell whether two amplicon ids have equal abundances.
int swb200_load_db_rows(swb200_ctx *c, const uint64_t *words, uint32_t stride_words, const uint32_t *len, const uint64_t *abundance,
                        uint32_t n_total, uint32_t first, uint32_t count, const uint32_t *run_start, uint32_t n_runs) {
  API_BEGIN(c)
  if (n_total == 0 || stride_words == 0 || count == 0 || static_cast<uint64_t>(first) + count > n_total || !words || !len || !abundance ||
      !run_start || n_runs == 0 || run_start[0] != 0 || run_start[n_runs] != n_total || n_total >= 0xFFFFFFF0u) {
    g_err = "load_db_rows: bad argument";
    return SWB200_EINVAL;
  }
  c->indexed = c->have_network = c->clustered = false;
  c->n = n_total; c->stride = stride_words; c->batch = 2; c->n_padded = n_total;
  c->db_sharded = true; c->row_first = first; c->row_count = count;
  c->words.alloc(static_cast<size_t>(count) * stride_words + 2);
  c->len.alloc(count);
  c->abundance.alloc(count);
  c->run_start_d.alloc(static_cast<size_t>(n_runs) + 1);
  c->n_runs = n_runs;
  c->tic();
  db_upload(c, c->words.p, words, static_cast<size_t>(count) * stride_words * 8);
  db_upload(c, c->len.p, len, static_cast<size_t>(count) * 4);
  db_upload(c, c->abundance.p, abundance, static_cast<size_t>(count) * 8);
  db_upload(c, c->run_start_d.p, run_start, (static_cast<size_t>(n_runs) + 1) * 4);
  return db_finalize(c);
  API_END()
}

int swb200_load_db_device(swb200_ctx *c, const void *d_words, uint32_t stride_words, const void *d_len, const void *d_abundance,
                          uint32_t n) {
  API_BEGIN(c)
  if (!d_words || !d_len || !d_abundance || n == 0 || stride_words == 0 || n >= 0xFFFFFFF0u) { g_err = "load_db_device: bad argument"; return SWB200_EINVAL; }
  db_alloc(c, n, stride_words);
  c->tic();
  CK(cudaMemcpyAsync(c->words.p, d_words, static_cast<size_t>(n) * stride_words * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->len.p, d_len, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->abundance.p, d_abundance, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToDevice, c->stream));
  return db_finalize(c);
  API_END()
}

int swb200_db_device(swb200_ctx *c, void **d_words, void **d_len, void **d_abundance) {
  if (!c || c->n == 0) { g_err = "db_device: no database"; return SWB200_EINVAL; }
  if (d_words) *d_words = c->words.p;
  if (d_len) *d_len = c->len.p;
  if (d_abundance) *d_abundance = c->abundance.p;
  return SWB200_OK;
}

int swb200_db_commit(swb200_ctx *c) {
  API_BEGIN(c)
  if (c->n == 0) { g_err = "db_commit: no database"; return SWB200_EINVAL; }
  c->db_pending = false;
  c->tic();
  return db_finalize(c);
  API_END()
}

static TileJoinParams tile_params(swb200_ctx *c) {
  TileJoinParams J{};
  J.words = c->words.p; J.len = c->len.p; J.abundance = c->abundance.p; J.n = c->n; J.stride = c->stride; J.K = c->jK;
  J.n_tiles = c->tj_tiles; J.t_lo = c->tj_lo; J.t_hi = c->tj_hi; J.cmax = c->tj_cmax;
  uint32_t idb = 12;
  while (idb < 32 && (1ull << idb) < c->n) ++idb;
  J.id_bits = idb; J.sorted_desc = c->sorted_desc ? 1 : 0;
  J.tile_count = c->tj_count.p; J.tile_off = c->tj_off.p; J.tile_cursor = c->tj_cursor.p;
  J.big_tiles = c->tj_big.p; J.big_count = reinterpret_cast<uint32_t *>(c->counters.p + 16);
  J.entries = c->tj_entries.p;
  J.edges = c->edges.p; J.edge_count = c->counters.p; J.edge_cap = c->edges.n;
  J.ncb = c->ncb; J.dup_flag = reinterpret_cast<uint32_t *>(c->counters.p + 8);
  J.stats = c->collect_stats ? c->counters.p + 1 : nullptr;
  return J;
}

// ---- tile store (d1_tilestore.cuh) ----------------------------------------------------------------------------------
static TileStoreParams ts_params(swb200_ctx *c) {
  TileStoreParams J{};
  J.words = c->words.p; J.len = c->len.p; J.abundance = c->abundance.p;
  J.ab_all = c->db_sharded ? nullptr : c->abundance.p; J.run_start = c->run_start_d.p; J.n_runs = c->n_runs;
  J.n = c->n; J.row_first = c->row_first; J.row_count = c->row_count;
  J.stride = c->stride; J.K = c->jK;
  J.kmask1 = c->jK >= 64 ? ~0ull : (c->jK > 32 ? (1ull << (2 * (c->jK - 32))) - 1 : 0ull);
  J.kmask0 = c->jK >= 32 ? ~0ull : (1ull << (2 * c->jK)) - 1;
  uint32_t idb = 12;
  while (idb < 32 && (1ull << idb) < c->n) ++idb;
  J.id_bits = idb; J.sorted_desc = c->sorted_desc ? 1 : 0; J.ncb = c->ncb;
  J.n_tiles = c->ts_tiles; J.t_lo = c->ts_lo; J.t_hi = c->ts_hi;
  // tile record: 8-byte entry (slim: rows gathered from the database), entry + packed row (fat), or — sharded database — entry + a
  // reference to the row where it lies: in the inbox it arrived in, or among this rank's own rows
  J.cap = c->ts_cap; J.rec_words = c->db_sharded ? 2 : (c->ts_fat ? c->stride + 1 : 1);
  J.row_base = c->db_sharded && c->dist_world ? reinterpret_cast<const unsigned long long *>(c->dist_peer[c->dist_rank] + kDistCtlBytes) : nullptr;
  J.q_cap = c->ts_qcap; J.out_cap = c->ts_outcap;
  J.store = c->ts_store.p; J.cursor = c->ts_cursor.p; J.ovf_head = c->ts_cursor.p + (c->ts_hi - c->ts_lo);
  J.ovf = c->ts_ovf.p; J.ovf_count = c->counters.p + 32; J.ovf_cap = c->ts_ovf_cap;
  J.ovf_abort = reinterpret_cast<uint32_t *>(c->counters.p + 34);
  // dense data: past 32 pair tests per amplicon (and 10^8 in all) the linear enumeration wins — single GPU with the
  // per-position tables on chip only; the tile_cmax test hook forces everything through the overflow path instead
  J.ovf_budget = (c->skew_fallback && c->tj_cmax_override < 2 && !c->too_long)
                     ? (static_cast<uint64_t>(c->n) * 16 + 50000000ull) : 0;
  J.edges = c->edges.p; J.edge_count = c->counters.p; J.edge_cap = c->edges.n;
  J.dup_flag = reinterpret_cast<uint32_t *>(c->counters.p + 8);
  J.stats = c->collect_stats ? c->counters.p + 1 : nullptr;
  J.appended = c->counters.p + 36;
  return J;
}

// tile geometry + every buffer of the tile-store path (no stream work: swb200_d1_reserve calls it ahead of time)
static uint32_t ts_prepare(swb200_ctx *c) {
  const uint32_t rw = c->db_sharded ? 2 : (c->ts_fat ? c->stride + 1 : 1);
  const uint32_t rec_bytes = 8 * (c->stride + 1) + (c->db_sharded ? 8 : 0);      // shared memory per record either way: entry (+ reference) + row
  uint32_t cap = std::min<uint32_t>(512, (36u * 1024u) / rec_bytes) & ~1u;      // 512 x 48 B: five CTAs of the join per SM (measured best, profiles/r2h)
  cap = std::max<uint32_t>(cap, 64);
  if (c->ts_cap_opt >= 64) cap = std::min(cap, c->ts_cap_opt & ~1u);
  const uint32_t fill = cap * 2 / 3;                       // mean records per tile; the slack absorbs the tail, the rest overflows (k_ts_big)
  const uint64_t want_tiles = (static_cast<uint64_t>(c->n) * 2 + fill - 1) / fill;
  c->ts_tiles = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(want_tiles, 0x7FFFFFFFull)));
  if (c->tj_cmax_override >= 2) cap = std::max<uint32_t>(2, std::min(cap, c->tj_cmax_override) & ~1u);
  c->ts_cap = cap;
  const uint32_t own_world = c->dist_world > 1 ? c->dist_world : static_cast<uint32_t>(c->shard_world);      // after dist_setup the job's ranks share the tiles
  const uint32_t own_rank = c->dist_world > 1 ? c->dist_rank : static_cast<uint32_t>(c->shard_rank);
  const uint32_t per = (c->ts_tiles + own_world - 1) / own_world;
  c->ts_lo = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(per) * own_rank, c->ts_tiles));
  c->ts_hi = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(c->ts_lo) + per, c->ts_tiles));
  const uint32_t T = c->ts_hi - c->ts_lo;
  // occupancy of the join: the kernel is bound by instruction issue and by shared-memory latency, so more resident warps pay;
  // 4 CTAs/SM leave 64 registers per thread, 6 leave 40 (ptxas: no spills at 5, 4 bytes at 6) when the tile fits 1/6 of the SM
  auto smem_for = [&](uint32_t q, uint32_t o) {
    return static_cast<size_t>(cap) * rec_bytes + (kTsBuckets + 2) * 4 + static_cast<size_t>(cap) * 4 + q * 4 + o * 8 + ((static_cast<size_t>(cap) * 2 + 15) & ~size_t(15));
  };
  const size_t sm_bytes = 233472;
  c->ts_occ = 4; c->ts_qcap = 1536; c->ts_outcap = 512;
  for (int occ = 6; occ > 4; --occ) {
    if (c->ts_occ_opt && occ > c->ts_occ_opt) continue;
    if ((smem_for(1024, 256) + 1024 + 256) * occ <= sm_bytes) { c->ts_occ = occ; c->ts_qcap = 1024; c->ts_outcap = 256; break; }
  }
  if (c->ts_occ_opt == 4) { c->ts_occ = 4; c->ts_qcap = 1536; c->ts_outcap = 512; }
  c->ts_smem = smem_for(c->ts_qcap, c->ts_outcap);
  c->ts_store.alloc(std::max<uint64_t>(static_cast<uint64_t>(T) * cap * rw, 2) + 2);
  c->ts_cursor.alloc(static_cast<size_t>(T) * 2 + 2);
  if (c->ts_ovf_cap == 0 || c->tj_cmax_override >= 2)
    c->ts_ovf_cap = std::max<uint64_t>(c->ts_ovf_cap, c->tj_cmax_override >= 2 ? static_cast<uint64_t>(c->n) * 2 / own_world + (1u << 16)
                                                                                : static_cast<uint64_t>(c->n) / 8 / own_world + (1u << 16));
  c->ts_ovf.alloc(c->ts_ovf_cap * (rw + 1));
  c->ts_route_cnt.alloc(kDistMaxWorld + 4);
  if (c->edges.n == 0) c->edges.alloc(std::max<size_t>(static_cast<size_t>(c->n) * 4 / own_world + (1u << 16), 1u << 16));
  return per;
}

// "Hashing sequences" for the tile store: ONE scatter pass (or, multi-GPU, the index exchange).  No host synchronisation.
static void index_tilestore(swb200_ctx *c) {
  const uint32_t per = ts_prepare(c);
  const uint32_t T = c->ts_hi - c->ts_lo;
  CK(cudaMemsetAsync(c->counters.p, 0, 17 * 8, c->stream));
  CK(cudaMemsetAsync(c->counters.p + 32, 0, 5 * 8, c->stream));
  CK(cudaMemsetAsync(c->ts_cursor.p, 0, static_cast<size_t>(T) * 4, c->stream));
  CK(cudaMemsetAsync(c->ts_cursor.p + T, 0xFF, static_cast<size_t>(T) * 4, c->stream));
  if (c->exchange()) {
    // multi-GPU: hash only this rank's rows, route every record to the owner of its tile over NVLink (d1_tsroute.cuh)
    TsRouteParams R{};
    R.J = ts_params(c);
    R.rank = c->dist_rank; R.world = c->dist_world;
    R.tiles_per_rank = per;
    for (uint32_t r = 0; r < R.world; ++r) R.peer[r] = c->dist_peer[r];
    R.rec_words = c->ts_fat ? c->stride + 1 : 2;
    R.inbox_cap = (c->dist_buffer_bytes - kDistCtlBytes) / (static_cast<uint64_t>(R.world) * R.rec_words * 8);
    R.counters = c->ts_route_cnt.p;
    R.done_ctas = reinterpret_cast<uint32_t *>(c->ts_route_cnt.p + kDistMaxWorld);
    R.err = R.done_ctas + 4;
    R.epoch = ++c->idx_epoch;
    if (R.epoch == 1) CK(cudaMemsetAsync(c->ts_route_cnt.p, 0, (kDistMaxWorld + 4) * 8, c->stream));
    else CK(cudaMemsetAsync(c->ts_route_cnt.p, 0, kDistMaxWorld * 8, c->stream));
    if (!c->db_sharded) {                                      // replicated database: this rank hashes an equal share of the rows
      const uint32_t share = ((c->n + R.world - 1) / R.world + 1u) & ~1u;          // even: the TMA row staging needs 16-byte aligned sources
      R.J.row_first = std::min<uint64_t>(static_cast<uint64_t>(share) * R.rank, c->n);
      R.J.row_count = std::min<uint32_t>(share, c->n - R.J.row_first);
      R.J.words += static_cast<size_t>(R.J.row_first) * c->stride;
      R.J.len += R.J.row_first;
      R.J.abundance += R.J.row_first;
    }
    const size_t smem = static_cast<size_t>(kTsRows) * c->stride * 8 + 16 + static_cast<size_t>(2 * kTsRows) * R.rec_words * 8;
    const unsigned route_grid = (R.J.row_count + kTsRows - 1) / kTsRows;
    const unsigned scat_grid = static_cast<unsigned>(c->sm_count * 8);
    k_ts_wait<<<1, 32, 0, c->stream>>>(R, 0);
    if (route_grid) {
      if (c->ts_fat) {
        CK(cudaFuncSetAttribute(k_ts_route<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        k_ts_route<true><<<route_grid, kTsRows, smem, c->stream>>>(R);
      } else {
        CK(cudaFuncSetAttribute(k_ts_route<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        k_ts_route<false><<<route_grid, kTsRows, smem, c->stream>>>(R);
      }
    }
    k_ts_wait<<<1, 32, 0, c->stream>>>(R, 1);
    R.J = ts_params(c);                                        // the receiving side works on the local tiles (and, slim, on the whole database)
    if (c->ts_fat) k_ts_scatter_inbox<true><<<scat_grid, 256, 0, c->stream>>>(R);
    else k_ts_scatter_inbox<false><<<scat_grid, 256, 0, c->stream>>>(R);
    CK(cudaGetLastError());
    c->launches += 4;
  } else if (T) {
    TileStoreParams J = ts_params(c);
    const size_t smem = static_cast<size_t>(kTsRows) * c->stride * 8 + 16;
    CK(cudaFuncSetAttribute(k_ts_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k_ts_scatter<<<(c->row_count + kTsRows - 1) / kTsRows, kTsRows, smem, c->stream>>>(J);
    CK(cudaGetLastError());
    c->launches += 1;
  }
}

// "Hashing sequences" for the enumeration kernels: Zobrist hash of every amplicon, open-addressing table, blocked
// Bloom filter, duplicate check.  Returns true when two amplicons are identical.
static bool index_table(swb200_ctx *c) {
  c->n_slots = table_slots(c->n);
  c->n_filter_blocks = std::max<uint64_t>(1, c->n_slots * c->bloom_bytes_per_slot / 8);
  if (c->n_filter_blocks > (1ull << 32)) c->n_filter_blocks = 1ull << 32;
  c->slots.alloc(c->n_slots);
  c->filter.alloc(c->n_filter_blocks);
  c->hashes.alloc(c->n);
  CK(cudaMemsetAsync(c->slots.p, 0xFF, c->n_slots * sizeof(Slot), c->stream));
  CK(cudaMemsetAsync(c->filter.p, 0, c->n_filter_blocks * 8, c->stream));
  CK(cudaMemsetAsync(c->counters.p, 0, 16 * 8, c->stream));
  D1Params P = c->params();
  const size_t zbytes = static_cast<size_t>(c->zlen) * 32;
  CK(cudaFuncSetAttribute(k_d1_index, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(zbytes)));
  const int grid = std::min<int>(c->sm_count * 8, (c->n + 255) / 256);
  k_d1_index<<<grid, 256, zbytes, c->stream>>>(P);
  CK(cudaGetLastError());
  k_d1_dupcheck<<<(c->n + 255) / 256, 256, 0, c->stream>>>(P);
  CK(cudaGetLastError());
  c->launches += 2;
  uint32_t dup = 0;
  CK(cudaMemcpyAsync(&dup, c->counters.p + 8, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return dup != 0;
}

int swb200_d1_index(swb200_ctx *c) {
  API_BEGIN(c)
  if (c->n == 0) { g_err = "d1_index: no database loaded"; return SWB200_EINVAL; }
  if (c->db_pending) { g_err = "d1_index: swb200_load_db_shard must be followed by the row exchange and swb200_db_commit"; return SWB200_EINVAL; }
  c->join_active = c->tile_active = c->ts_active = false;
  c->jK = std::min<uint32_t>(64, c->min_len / 2);
  const bool join = c->enum_mode == SWB200_ENUM_JOIN && c->jK >= 8;
  if (c->db_sharded) {
    if (!(join && c->join_kernel == 0 && c->stride <= 64 && c->max_len < 8192 && c->dist_world > 1 && c->ts_fat && c->unsorted == 0)) {
      g_err = "d1_index: a sharded database (swb200_load_db_rows) needs swb200_dist_setup, tile_rows = 1, the default join, a database sorted "
              "by abundance and sequences of 16..8191 nt";
      return SWB200_EUNSUPPORTED;
    }
  }
  if (join && c->join_kernel == 0 && c->stride <= 64 && c->max_len < 8192) {
    // JOIN over the tile store (d1_tilestore.cuh): one scatter pass, fat records in fixed-capacity tile slots
    c->tic();
    index_tilestore(c);
    c->toc(1);
    c->indexed = true;
    c->join_active = c->ts_active = true;
    c->have_network = c->clustered = false;
    return SWB200_OK;
  }
  if (join && c->join_kernel == 2 && c->stride <= 32 && c->max_len < 8192) {
    // JOIN, partitioned: two K-mer entries per amplicon appended to hash-range tiles (d1_tilejoin.cuh)
    c->tj_cmax = c->stride <= 6 ? 768 : (c->stride <= 14 ? 384 : 192);
    const uint64_t want_tiles = (static_cast<uint64_t>(c->n) * 2 + c->tj_cmax / 2 - 1) / (c->tj_cmax / 2);
    c->tj_tiles = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(want_tiles, 0x7FFFFFFFull)));
    const uint32_t per = (c->tj_tiles + c->shard_world - 1) / c->shard_world;
    c->tj_lo = std::min<uint64_t>(static_cast<uint64_t>(per) * c->shard_rank, c->tj_tiles);
    c->tj_hi = std::min<uint64_t>(static_cast<uint64_t>(c->tj_lo) + per, c->tj_tiles);
    const uint32_t T = c->tj_hi - c->tj_lo;
    if (c->tj_cmax_override >= 2) c->tj_cmax = std::min(c->tj_cmax, c->tj_cmax_override);
    c->tj_smem = static_cast<size_t>(c->tj_cmax) * 8 + static_cast<size_t>(c->tj_cmax) * c->stride * 8 + kTjOutCap * 8 +
                 (kTjBuckets + 2) * 4 + (static_cast<size_t>(c->tj_cmax) + 2) * 4;
    c->tj_count.alloc(T + 1); c->tj_cursor.alloc(T + 1); c->tj_big.alloc(T + 1); c->tj_off.alloc(T + 2);
    c->tic();
    CK(cudaMemsetAsync(c->counters.p, 0, 17 * 8, c->stream));
    CK(cudaMemsetAsync(c->tj_count.p, 0, static_cast<size_t>(T + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->tj_cursor.p, 0, static_cast<size_t>(T + 1) * 4, c->stream));
    TileJoinParams J = tile_params(c);
    const unsigned pb = (c->n + 255) / 256;
    unsigned long long total = 0;
    if (T) {
      const bool use_list = c->shard_world > 1;          // hash every amplicon once, keep the pieces of this rank's tiles in a list
      unsigned long long listed[kPlistSubs * kPlistPad];
      if (use_list) {
        // capacity of a sub-list: its share of the rank's expected pieces + 25 % + slack
        const uint64_t cap = (static_cast<uint64_t>(c->n) * 2 / c->shard_world * 5 / 4 + (1u << 16)) / kPlistSubs + 1024;
        c->tj_plist_ent.alloc(cap * kPlistSubs); c->tj_plist_tile.alloc(cap * kPlistSubs); c->tj_plist_cnt.alloc(kPlistSubs * kPlistPad);
        J.plist_ent = c->tj_plist_ent.p; J.plist_tile = c->tj_plist_tile.p; J.plist_cap = cap;
        J.plist_n = c->tj_plist_cnt.p;
        CK(cudaMemsetAsync(c->tj_plist_cnt.p, 0, sizeof listed, c->stream));
        k_tile_partition_list<<<pb, 256, 0, c->stream>>>(J);
      } else {
        k_tile_partition<true><<<pb, 256, 0, c->stream>>>(J);
      }
      k_tile_scan<<<1, 1024, 0, c->stream>>>(J);
      CK(cudaMemcpyAsync(&total, c->tj_off.p + T, 8, cudaMemcpyDeviceToHost, c->stream));
      if (use_list) CK(cudaMemcpyAsync(listed, c->tj_plist_cnt.p, sizeof listed, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      c->tj_total = total;
      c->tj_entries.alloc(std::max<unsigned long long>(total, 1));
      J.entries = c->tj_entries.p;
      unsigned long long longest_list = 0;
      if (use_list) for (uint32_t s = 0; s < kPlistSubs; ++s) longest_list = std::max(longest_list, listed[s * kPlistPad]);
      if (use_list && longest_list <= J.plist_cap) {
        if (longest_list) k_tile_scatter_list<<<dim3(static_cast<unsigned>((longest_list + 255) / 256), kPlistSubs), 256, 0, c->stream>>>(J);
      } else {
        k_tile_partition<false><<<pb, 256, 0, c->stream>>>(J);      // single GPU, or a sub-list overflowed (a very skewed hash range)
      }
      CK(cudaGetLastError());
      c->launches += 3;
    }
    c->toc(1);
    c->indexed = true;
    c->join_active = c->tile_active = true;
    c->have_network = c->clustered = false;
    return SWB200_OK;
  }
  if (c->enum_mode == SWB200_ENUM_JOIN && c->jK >= 8) {
    // JOIN: the index is a multimap of two K-mer pieces per amplicon (d1_join.cuh); no Zobrist table, no filter
    const uint64_t slots = std::max<uint64_t>(64, (static_cast<uint64_t>(c->n) * 2 * 5 / 2 + 3) / 4 * 4);
    c->jtab_buckets = slots / 4;
    // multi-GPU: this rank owns the bucket (hash) range [b_lo, b_hi) of the table
    const uint64_t per = (c->jtab_buckets + c->shard_world - 1) / c->shard_world;
    c->jb_lo = std::min<uint64_t>(per * c->shard_rank, c->jtab_buckets);
    c->jb_hi = std::min<uint64_t>(c->jb_lo + per, c->jtab_buckets);
    const uint64_t my_slots = std::max<uint64_t>(4, (c->jb_hi - c->jb_lo) * 4);
    c->jtab.alloc(my_slots);
    c->tic();
    CK(cudaMemsetAsync(c->jtab.p, 0xFF, my_slots * 8, c->stream));
    CK(cudaMemsetAsync(c->counters.p, 0, 16 * 8, c->stream));
    NetJoinParams J{};
    J.words = c->words.p; J.len = c->len.p; J.abundance = c->abundance.p; J.n = c->n; J.stride = c->stride; J.K = c->jK;
    J.table = c->jtab.p; J.n_buckets = c->jtab_buckets; J.b_lo = c->jb_lo; J.b_hi = c->jb_hi;
    k_join_index<<<static_cast<unsigned>((static_cast<uint64_t>(c->n) * 2 + 255) / 256), 256, 0, c->stream>>>(J);
    CK(cudaGetLastError());
    c->launches += 1;
    c->toc(1);
    c->indexed = true;
    c->join_active = true;
    c->have_network = c->clustered = false;
    return SWB200_OK;
  }
  if (c->too_long) { g_err = "sequences longer than 5,000 nt need enum_mode JOIN (the enumeration kernels keep per-position tables on chip)"; return SWB200_EUNSUPPORTED; }
  c->tic();
  const bool dup = index_table(c);
  c->toc(1);
  c->indexed = true;
  c->have_network = c->clustered = false;
  if (dup) { g_err = "some fasta entries have identical sequences"; return SWB200_EDUPLICATE; }
  API_END()
}

static void run_network(swb200_ctx *c) {
  D1Params P = c->params();
  // this context's share of the seeds: contiguous, batch-aligned range
  const uint64_t per = (static_cast<uint64_t>(c->n_padded / c->batch) + c->shard_world - 1) / c->shard_world;
  const uint64_t b0 = std::min<uint64_t>(per * c->shard_rank, c->n_padded / c->batch);
  const uint64_t b1 = std::min<uint64_t>(b0 + per, c->n_padded / c->batch);
  P.seed_begin = static_cast<uint32_t>(b0 * c->batch);
  P.seed_end = static_cast<uint32_t>(std::min<uint64_t>(b1 * c->batch, c->n));
  if (P.seed_begin > P.seed_end) P.seed_begin = P.seed_end;
  auto launch = [&](auto kernel, size_t smem) {
    int occ = 1;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kWarpsPerCta * 32, smem));
    occ = std::max(occ, 1);
    kernel<<<c->sm_count * occ, kWarpsPerCta * 32, smem, c->stream>>>(P);
  };
  const bool st = c->collect_stats != 0;
  const size_t smem1 = c->network_smem();
  const size_t smem2 = static_cast<size_t>(c->zlen) * kTStride + static_cast<size_t>(kWarpsPerCta) * 2 * c->batch * c->stride * 8 +
                       static_cast<size_t>(kWarpsPerCta) * sizeof(WarpScratch2) + kWarpsPerCta * 2 * 8;
  const bool v2_ok = c->enum_mode != SWB200_ENUM_FULL && static_cast<size_t>(c->zlen) * kTStride <= 60 * 1024 && c->max_len <= 990;
  const bool use_v2 = v2_ok && c->net_kernel != 1;
  if (use_v2) { if (st) launch(k_d1_network_half<true>, smem2); else launch(k_d1_network_half<false>, smem2); }
  else if (c->enum_mode == SWB200_ENUM_FULL) { if (st) launch(k_d1_network<0, true>, smem1); else launch(k_d1_network<0, false>, smem1); }
  else { if (st) launch(k_d1_network<1, true>, smem1); else launch(k_d1_network<1, false>, smem1); }
  CK(cudaGetLastError());
  c->launches += 1;
}

int swb200_d1_network(swb200_ctx *c, int no_cluster_breaking, uint64_t *n_links) {
  API_BEGIN(c)
  if (!c->indexed) { g_err = "d1_network: call swb200_d1_index first"; return SWB200_EINVAL; }
  c->ncb = no_cluster_breaking ? 1 : 0;
  if (c->edges.n == 0) c->edges.alloc(std::max<size_t>(static_cast<size_t>(c->n) * 4, 1u << 16));
  c->tic();
  for (int attempt = 0; attempt < 4; ++attempt) {
    CK(cudaMemsetAsync(c->counters.p, 0, 10 * 8, c->stream));
    if (c->ts_active) {
      TileStoreParams J = ts_params(c);
      const uint32_t T = c->ts_hi - c->ts_lo;
      if (T) {
        auto launch = [&](auto join, auto big) {
          CK(cudaFuncSetAttribute(join, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->ts_smem)));
          join<<<T, 256, c->ts_smem, c->stream>>>(J);
          big<<<c->sm_count * 4, 256, 0, c->stream>>>(J);
        };
        auto pick = [&](auto occ) {
          constexpr int O = decltype(occ)::value;
          if (c->ts_fat && !c->db_sharded) { if (c->collect_stats) launch(k_ts_join<true, true, O>, k_ts_big<true, true>); else launch(k_ts_join<true, false, O>, k_ts_big<true, false>); }
          else { if (c->collect_stats) launch(k_ts_join<false, true, O>, k_ts_big<false, true>); else launch(k_ts_join<false, false, O>, k_ts_big<false, false>); }
        };
        if (c->ts_occ >= 6) pick(std::integral_constant<int, 6>{});
        else if (c->ts_occ == 5) pick(std::integral_constant<int, 5>{});
        else pick(std::integral_constant<int, 4>{});
        if (c->db_sharded) {                       // the rows were read out of the inbox: the senders may overwrite it (next epoch)
          TsRouteParams R{};
          R.rank = c->dist_rank; R.world = c->dist_world; R.epoch = c->idx_epoch;
          for (uint32_t r = 0; r < R.world; ++r) R.peer[r] = c->dist_peer[r];
          k_ts_release<<<1, 32, 0, c->stream>>>(R);
          c->launches += 1;
        }
        c->launches += 2;
        CK(cudaGetLastError());
      }
    } else if (c->tile_active) {
      TileJoinParams J = tile_params(c);
      const uint32_t T = c->tj_hi - c->tj_lo;
      if (T) {
        if (c->collect_stats) {
          CK(cudaFuncSetAttribute(k_tile_join<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->tj_smem)));
          k_tile_join<true><<<T, 256, c->tj_smem, c->stream>>>(J);
          k_tile_join_big<true><<<c->sm_count * 4, 256, 0, c->stream>>>(J);
        } else {
          CK(cudaFuncSetAttribute(k_tile_join<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->tj_smem)));
          k_tile_join<false><<<T, 256, c->tj_smem, c->stream>>>(J);
          k_tile_join_big<false><<<c->sm_count * 4, 256, 0, c->stream>>>(J);
        }
        c->launches += 2;
        CK(cudaGetLastError());
      }
    } else if (c->join_active) {
      NetJoinParams J{};
      J.words = c->words.p; J.len = c->len.p; J.abundance = c->abundance.p; J.n = c->n; J.stride = c->stride; J.K = c->jK;
      J.table = c->jtab.p; J.n_buckets = c->jtab_buckets; J.b_lo = c->jb_lo; J.b_hi = c->jb_hi;
      J.edges = c->edges.p; J.edge_count = c->counters.p; J.edge_cap = c->edges.n;
      J.ncb = c->ncb; J.dup_flag = reinterpret_cast<uint32_t *>(c->counters.p + 8);
      J.stats = c->collect_stats ? c->counters.p + 1 : nullptr;
      J.seed_begin = 0;                 // every rank scans all lookups and walks only its own hash range
      J.seed_end = c->n;
      const uint64_t threads = static_cast<uint64_t>(J.seed_end - J.seed_begin) * 2;
      if (c->cands.n < static_cast<size_t>(c->n) * 6) c->cands.alloc(std::max<size_t>(static_cast<size_t>(c->n) * 6, 1u << 20));
      unsigned long long ncand = 0;
      for (int tries = 0; tries < 2 && threads; ++tries) {
        J.cands = c->cands.p; J.cand_cap = c->cands.n; J.cand_count = c->counters.p + 9;
        CK(cudaMemsetAsync(c->counters.p + 9, 0, 8, c->stream));
        if (tries) CK(cudaMemsetAsync(c->counters.p + 1, 0, 4 * 8, c->stream));
        k_join_candidates<<<static_cast<unsigned>(std::min<uint64_t>((threads + 255) / 256, static_cast<uint64_t>(c->sm_count) * 8)), 256, 0, c->stream>>>(J);
        c->launches += 1;
        CK(cudaMemcpyAsync(&ncand, c->counters.p + 9, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (ncand <= c->cands.n) break;
        c->cands.alloc(ncand + ncand / 8);
      }
      if (ncand) {
        k_join_verify<<<static_cast<unsigned>(std::min<uint64_t>((ncand + 255) / 256, static_cast<uint64_t>(c->sm_count) * 8)), 256, 0, c->stream>>>(J, ncand);
        c->launches += 1;
      }
      CK(cudaGetLastError());
    } else
    run_network(c);
    unsigned long long *host = static_cast<unsigned long long *>(c->staging(4096));   // one read-back (pinned): links, stats, duplicate flag, overflow counters
    CK(cudaMemcpyAsync(host, c->counters.p, 37 * 8, cudaMemcpyDeviceToHost, c->stream));
    const bool exchanged = c->ts_active && c->exchange();
    if (exchanged) CK(cudaMemcpyAsync(host + 40, reinterpret_cast<uint32_t *>(c->ts_route_cnt.p + kDistMaxWorld) + 4, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->n_edges = host[0];
    for (int i = 0; i < 4; ++i) c->stats[i] = host[1 + i];
    c->stats[4] = c->n_edges;
    c->stats[6] = host[5];
    if (c->tile_active && c->collect_stats) c->stats[0] = c->tj_total;     // entries of this rank's tiles (2 per amplicon over all ranks)
    if (c->ts_active) {
      const unsigned long long *ov = host + 32;                            // [0] overflow records [1] their pair tests [2] abort flag
      c->ts_overflow = ov[0];
      if (c->collect_stats) c->stats[0] = host[36];                        // records appended to this rank's tiles (2 per amplicon over all ranks)
      if (ov[0] > c->ts_ovf_cap) {                                         // the overflow list itself overflowed: grow it, index again
        c->ts_ovf_cap = ov[0] + ov[0] / 8 + 1024;
        index_tilestore(c);
        continue;
      }
      if (static_cast<uint32_t>(ov[2])) {
        // dense data: a few K-mers are shared by so many amplicons that the pairwise sweep of their tiles is quadratic;
        // the enumeration kernels are linear in the data, like the reference (src/algod1.cc:606-670)
        c->ts_active = c->join_active = false;
        c->ts_fallbacks++;
        if (index_table(c)) { c->toc(2); g_err = "some fasta entries have identical sequences"; return SWB200_EDUPLICATE; }
        continue;
      }
    }
    if (exchanged) {
      const uint32_t *rerr = reinterpret_cast<const uint32_t *>(host + 40);
      if (rerr[0]) { c->toc(2); g_err = "d1_index: a peer did not take part in the index exchange within 5 s"; return SWB200_ECUDA; }
      if (rerr[1]) { c->toc(2); g_err = "d1_index: an index inbox overflowed; set up larger peer buffers (swb200_dist_buffer_bytes)"; return SWB200_ENOMEM; }
    }
    if (c->join_active && static_cast<uint32_t>(host[8])) { c->toc(2); g_err = "some fasta entries have identical sequences"; return SWB200_EDUPLICATE; }
    if (c->n_edges <= c->edges.n) break;
    c->edges.alloc(c->n_edges + c->n_edges / 8);    // link list overflowed: grow and redo (dense data)
  }
  c->toc(2);
  c->have_network = true;
  c->clustered = false;
  if (n_links) *n_links = c->n_edges;
  API_END()
}

int swb200_d1_export_links(swb200_ctx *c, uint32_t *pairs) {
  API_BEGIN(c)
  if (!c->have_network || !pairs) { g_err = "export_links: no network / null buffer"; return SWB200_EINVAL; }
  CK(cudaMemcpyAsync(pairs, c->edges.p, c->n_edges * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END()
}

int swb200_d1_import_links(swb200_ctx *c, const uint32_t *pairs, uint64_t n_links) {
  API_BEGIN(c)
  if (!c->indexed || (!pairs && n_links)) { g_err = "import_links: bad state / null buffer"; return SWB200_EINVAL; }
  c->edges.alloc(std::max<uint64_t>(n_links, 1));
  CK(cudaMemcpyAsync(c->edges.p, pairs, n_links * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->n_edges = n_links;
  c->have_network = true;
  c->clustered = false;
  API_END()
}

int swb200_d1_links_device(swb200_ctx *c, void **d_pairs, uint64_t *n_links) {
  if (!c || !c->have_network) { g_err = "links_device: no network"; return SWB200_EINVAL; }
  if (d_pairs) *d_pairs = c->edges.p;
  if (n_links) *n_links = c->n_edges;
  return SWB200_OK;
}

int swb200_d1_import_links_device(swb200_ctx *c, const void *d_pairs, uint64_t n_links) {
  API_BEGIN(c)
  if (!c->indexed || (!d_pairs && n_links)) { g_err = "import_links_device: bad state"; return SWB200_EINVAL; }
  if (d_pairs != c->edges.p) {
    // the gathered list replaces the local one; the buffer only ever grows (no cudaMalloc/cudaFree per step)
    if (c->edges.n < n_links) c->edges.alloc(n_links + n_links / 8);
    CK(cudaMemcpyAsync(c->edges.p, d_pairs, n_links * 8, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  c->n_edges = n_links;
  c->have_network = true;
  c->clustered = false;
  API_END()
}

int swb200_d1_get_network(swb200_ctx *c, uint64_t *row_ptr, uint32_t *col) {
  API_BEGIN(c)
  if (c->db_sharded) { g_err = "d1_get_network: not available on a sharded database (use the multi-GPU entry points)"; return SWB200_EUNSUPPORTED; }
  if (!c->have_network || !row_ptr) { g_err = "get_network: no network / null buffer"; return SWB200_EINVAL; }
  // output path (-j writer), not the timed hot path: counting sort by source on the host, rows ascending
  std::vector<uint2> e(c->n_edges);
  CK(cudaMemcpyAsync(e.data(), c->edges.p, c->n_edges * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::fill(row_ptr, row_ptr + c->n + 1, 0);
  for (const uint2 &x : e) row_ptr[x.x + 1]++;
  for (uint32_t i = 0; i < c->n; ++i) row_ptr[i + 1] += row_ptr[i];
  if (col) {
    std::vector<uint64_t> cur(row_ptr, row_ptr + c->n);
    for (const uint2 &x : e) col[cur[x.x]++] = x.y;
    for (uint32_t i = 0; i < c->n; ++i) std::sort(col + row_ptr[i], col + row_ptr[i + 1]);
  }
  API_END()
}

// geometry of the packed relaxation word swarm | generation | parent (d1_kernels.cuh: k_cluster_persistent<.., PACK>); false: unpacked
static bool cluster_pack_bits(const swb200_ctx *c, uint32_t &ib, uint32_t &gb) {
  ib = 1;
  while (ib < 32 && (1ull << ib) <= c->n) ++ib;              // 2^ib - 1 >= n: the all-ones parent field is no amplicon
  gb = 2 * ib <= 64 ? std::min<uint32_t>(32, 64 - 2 * ib) : 0;
  if (c->cluster_gen_bits) gb = std::min<uint32_t>(gb, c->cluster_gen_bits);
  return c->cluster_pack && gb >= (c->cluster_gen_bits ? 1u : 10u);
}

static void run_cluster(swb200_ctx *c) {
  const uint32_t n = c->n;
  const uint64_t m = c->n_edges;
  c->label.alloc(n); c->generation.alloc(n); c->parent.alloc(n); c->key.alloc(n);
  uint32_t *changed = reinterpret_cast<uint32_t *>(c->counters.p + 9);
  const int vb = (n + 255) / 256;
  const int eb = static_cast<int>(std::min<uint64_t>((m + 255) / 256, static_cast<uint64_t>(c->sm_count) * 16));
  uint32_t *h_changed = static_cast<uint32_t *>(c->staging(64));
  if (c->cluster_kernel == 6) {
    // links bucketed by source block, one persistent kernel (d1_bucket.cuh), world = 1: measured 1.39 ms at 10 M against the 1.08 ms of
    // k_cluster_persistent (the rounds are faster, 0.73 ms, but bucketing costs 0.33 ms once) — the multi-GPU kernel, an option here
    BucketParams B{};
    DistParams &D = B.D;
    D.rank = 0; D.world = 1; D.n = n;
    D.n_local = static_cast<uint32_t>((static_cast<uint64_t>(n) + kDistBlock - 1) / kDistBlock * kDistBlock);
    D.edges = c->edges.p; D.m_local = m;
    c->key.alloc(D.n_local);
    D.nwords = D.n_local / 32;
    c->cl_bits.alloc(static_cast<size_t>(D.nwords) * 3);
    B.nblk = D.n_local / kDistBlock;
    c->bk_flag.alloc(static_cast<size_t>(B.nblk) * 3 + 16); c->bk_count.alloc(B.nblk + 1); c->bk_off.alloc(static_cast<size_t>(B.nblk) + 2);
    if (c->bk_links.n < m) c->bk_links.alloc(std::max<uint64_t>(m + m / 8, 1));
    c->bk_unit.alloc(c->bk_links.n / kBkUnit + 2); c->bk_act.alloc(c->bk_links.n / kBkUnit + 2);
    D.key = c->key.p; D.parent = c->parent.p; D.label = c->label.p; D.generation = c->generation.p; D.bits = c->cl_bits.p;
    D.lcnt = c->counters.p + 44;                     // unused at world = 1
    D.epoch_base = 0;
    D.lflags = reinterpret_cast<uint32_t *>(c->counters.p + 22);
    D.gbar = reinterpret_cast<unsigned int *>(c->counters.p + 42);
    B.bcount = c->bk_count.p; B.boff = c->bk_off.p; B.blinks = c->bk_links.p; B.blinks_cap = c->bk_links.n;
    B.unit_blk = c->bk_unit.p; B.act_list = c->bk_act.p; B.unit_cap = c->bk_unit.n - 1; B.act_n = reinterpret_cast<uint32_t *>(c->counters.p + 46);
    if (std::getenv("SWB200_CLUSTER_TS")) {
      c->cl_ts.alloc(128);
      CK(cudaMemsetAsync(c->cl_ts.p, 0, 128 * 8, c->stream));
      D.ts = c->cl_ts.p;
    }
    CK(cudaMemsetAsync(c->counters.p + 42, 0, 2 * 8, c->stream));
    const size_t dyn = static_cast<size_t>(kDistChunk) * sizeof(DistRec);
    bool pack = cluster_pack_bits(c, B.ib, B.gb);
    uint32_t *h_deep = static_cast<uint32_t *>(c->staging(64)) + 8;
    for (;;) {
      const void *kern = pack ? reinterpret_cast<const void *>(k_cluster_bucket<true>) : reinterpret_cast<const void *>(k_cluster_bucket<false>);
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
      int occ = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, dyn));
      const uint64_t want = (std::max<uint64_t>(m, n) + 255) / 256;
      const unsigned grid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(want, static_cast<uint64_t>(c->sm_count) * std::max(occ, 1))));
      void *args[] = {&B};
      CK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), args, dyn, c->stream));
      c->launches++;
      if (!pack) break;
      CK(cudaMemcpyAsync(h_deep, D.lflags + 6, 4, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (!*h_deep) break;
      pack = false;                                            // a generation beyond gb bits: once more with 32-bit generations
      c->cluster_unpacked_reruns++;
    }
    CK(cudaMemcpyAsync(reinterpret_cast<uint32_t *>(c->counters.p + 19), reinterpret_cast<uint32_t *>(c->counters.p + 22) + 4, 4, cudaMemcpyDeviceToDevice, c->stream));   // rounds
    if (D.ts) {
      unsigned long long h[128];
      CK(cudaMemcpyAsync(h, c->cl_ts.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      std::fprintf(stderr, "[cluster_bucket n=%u m=%llu] us (init, count + scan, scatter + round 0, rounds ..., parents + unpack):", n, static_cast<unsigned long long>(m));
      for (int i = 1; i < 128 && h[i]; ++i) std::fprintf(stderr, " %.1f", (h[i] - h[i - 1]) * 1e-3);
      std::fprintf(stderr, "\n");
    }
    return;
  }
  if (c->cluster_kernel == 4) {
    // frontier relaxation over 8-slot out-rows, one persistent cooperative kernel (d1_frontier.cuh)
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cluster_frontier, 256, 0));
    const uint64_t want = (std::max<uint64_t>(m, n) + 255) / 256;
    const unsigned grid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(want, static_cast<uint64_t>(c->sm_count) * std::max(occ, 1))));
    FrontierParams F{};
    F.edges = c->edges.p; F.m = m; F.n = n;
    F.key = c->key.p; F.parent = c->parent.p; F.label = c->label.p; F.generation = c->generation.p;
    F.nwords = (n + 31) / 32;
    c->cl_bits.alloc(static_cast<size_t>(F.nwords) * 3);
    c->fr_deg.alloc(n); c->fr_adj.alloc(static_cast<size_t>(n) * kFrSlots);
    if (c->fr_spill.n < m) c->fr_spill.alloc(std::max<uint64_t>(m + m / 8, 1));      // worst case: one hub owns every link
    F.bits = c->cl_bits.p; F.deg = c->fr_deg.p; F.adj = c->fr_adj.p;
    F.spill = c->fr_spill.p; F.spill_n = c->counters.p + 35; F.spill_cap = c->fr_spill.n;
    F.flags = reinterpret_cast<uint32_t *>(c->counters.p + 17); F.rounds_out = reinterpret_cast<uint32_t *>(c->counters.p + 19);
    if (std::getenv("SWB200_CLUSTER_TS")) {                    // profiling aid: phase timestamps of the persistent kernel on stderr
      c->cl_ts.alloc(64);
      CK(cudaMemsetAsync(c->cl_ts.p, 0, 64 * 8, c->stream));
      F.ts = c->cl_ts.p;
    }
    void *args[] = {&F};
    CK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_cluster_frontier), dim3(grid), dim3(256), args, 0, c->stream));
    c->launches++;
    if (F.ts) {
      unsigned long long h[64];
      CK(cudaMemcpyAsync(h, c->cl_ts.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      std::fprintf(stderr, "[cluster_frontier n=%u m=%llu] us (init, build + round 0, rounds ..., parents + unpack):", n, static_cast<unsigned long long>(m));
      for (int i = 1; i < 64 && h[i]; ++i) std::fprintf(stderr, " %.1f", (h[i] - h[i - 1]) * 1e-3);
      std::fprintf(stderr, "\n");
    }
    return;
  }
  if (c->cluster_kernel == 3 && m < 0xFFFFFFF0ull) {
    // links counting-sorted by source + frontier relaxation, one persistent cooperative kernel (d1_cluster.cuh)
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cluster_csr, 256, 0));
    const uint64_t want = (std::max<uint64_t>(m, n) + 255) / 256;
    const unsigned grid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(want, static_cast<uint64_t>(c->sm_count) * std::max(occ, 1))));
    ClusterParams C{};
    C.edges = c->edges.p; C.m = m; C.n = n;
    C.key = c->key.p; C.parent = c->parent.p; C.label = c->label.p; C.generation = c->generation.p;
    C.nwords = (n + 31) / 32;
    c->cl_bits.alloc(static_cast<size_t>(C.nwords) * 3);
    c->cl_deg.alloc(n); c->cl_row.alloc(static_cast<size_t>(n) + 1);
    c->cl_srcs.alloc(std::max<uint64_t>(m, 1)); c->cl_dsts.alloc(std::max<uint64_t>(m, 1));
    c->cl_tot.alloc(grid);
    C.bits = c->cl_bits.p; C.deg = c->cl_deg.p; C.row = c->cl_row.p; C.srcs = c->cl_srcs.p; C.dsts = c->cl_dsts.p;
    C.cta_tot = c->cl_tot.p;
    C.flags = reinterpret_cast<uint32_t *>(c->counters.p + 17); C.rounds_out = reinterpret_cast<uint32_t *>(c->counters.p + 19);
    if (std::getenv("SWB200_CLUSTER_TS")) {                    // profiling aid: phase timestamps of the persistent kernel on stderr
      c->cl_ts.alloc(64);
      CK(cudaMemsetAsync(c->cl_ts.p, 0, 64 * 8, c->stream));
      C.ts = c->cl_ts.p;
    }
    void *args[] = {&C};
    CK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_cluster_csr), dim3(grid), dim3(256), args, 0, c->stream));
    c->launches++;
    if (C.ts) {
      unsigned long long h[64];
      CK(cudaMemcpyAsync(h, c->cl_ts.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      std::fprintf(stderr, "[cluster_csr n=%u m=%llu] us:", n, static_cast<unsigned long long>(m));
      for (int i = 1; i < 64 && h[i]; ++i) std::fprintf(stderr, " %.1f", (h[i] - h[i - 1]) * 1e-3);
      std::fprintf(stderr, "\n");
    }
    return;
  }
  if (c->cluster_kernel == 0 || c->cluster_kernel == 5 || c->cluster_kernel == 3) {
    // fused label+generation relaxation as one persistent cooperative kernel over the unsorted link list (d1_kernels.cuh: k_cluster_persistent)
    const uint2 *e_p = c->edges.p;
    uint64_t m_ = m;
    unsigned long long *key_p = c->key.p;
    uint32_t *par_p = c->parent.p, *lab_p = c->label.p, *gen_p = c->generation.p, n_ = n;
    uint32_t *flags_p = reinterpret_cast<uint32_t *>(c->counters.p + 17), *rounds_p = reinterpret_cast<uint32_t *>(c->counters.p + 19);
    uint32_t nwords = (n + 31) / 32;
    c->cl_bits.alloc(static_cast<size_t>(nwords) * 3);
    uint32_t *bits_p = c->cl_bits.p;
    // packed relaxation word swarm | generation | parent (one atomicMin settles the parent too, no parent pass) while the ids leave
    // >= 10 bits for the generation; a deeper swarm raises bit 31 of the round count and the unpacked kernel runs instead
    uint32_t ib, gb;
    bool pack = cluster_pack_bits(c, ib, gb);
    uint32_t *h_rounds = static_cast<uint32_t *>(c->staging(64)) + 8;
    for (;;) {
      const void *kern = pack ? (c->cluster_hints ? reinterpret_cast<const void *>(k_cluster_persistent<true, true>) : reinterpret_cast<const void *>(k_cluster_persistent<false, true>))
                              : (c->cluster_hints ? reinterpret_cast<const void *>(k_cluster_persistent<true, false>) : reinterpret_cast<const void *>(k_cluster_persistent<false, false>));
      int occ = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
      const uint64_t want = (std::max<uint64_t>(m, n) + 255) / 256;
      const unsigned grid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(want, static_cast<uint64_t>(c->sm_count) * std::max(occ, 1))));
      void *args[] = {&e_p, &m_, &key_p, &par_p, &lab_p, &gen_p, &n_, &flags_p, &rounds_p, &bits_p, &nwords, &ib, &gb};
      CK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), args, 0, c->stream));
      c->launches++;
      if (!pack) break;
      CK(cudaMemcpyAsync(h_rounds, rounds_p, 4, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (!(*h_rounds & 0x80000000u)) break;
      pack = false;                                            // a generation beyond gb bits: once more with 32-bit generations
      c->cluster_unpacked_reruns++;
    }
    return;
  }
  if (c->cluster_kernel != 1) {
    // fused label+generation relaxation, host-looped (d1_kernels.cuh: k_key_*)
    k_key_init<<<vb, 256, 0, c->stream>>>(c->key.p, c->parent.p, n);
    c->launches++;
    for (int round = 0; m > 0 && round < 1 << 20; ++round) {
      CK(cudaMemsetAsync(changed, 0, 4, c->stream));
      k_key_relax<<<std::max(eb, 1), 256, 0, c->stream>>>(c->edges.p, m, c->key.p, changed);
      c->launches++;
      CK(cudaMemcpyAsync(h_changed, changed, 4, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (*h_changed == 0) break;
    }
    if (m > 0) { k_key_parent<<<std::max(eb, 1), 256, 0, c->stream>>>(c->edges.p, m, c->key.p, c->parent.p); c->launches++; }
    k_key_unpack<<<vb, 256, 0, c->stream>>>(c->key.p, c->label.p, c->generation.p, n);
    c->launches++;
    CK(cudaGetLastError());
    return;
  }
  k_label_init<<<vb, 256, 0, c->stream>>>(c->label.p, n);
  c->launches++;
  for (int round = 0; m > 0 && round < 1 << 20; ++round) {
    CK(cudaMemsetAsync(changed, 0, 4, c->stream));
    k_label_edges<<<std::max(eb, 1), 256, 0, c->stream>>>(c->edges.p, m, c->label.p, changed);
    k_label_jump<<<vb, 256, 0, c->stream>>>(c->label.p, n, changed);
    c->launches += 2;
    CK(cudaMemcpyAsync(h_changed, changed, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (*h_changed == 0) break;
  }
  k_bfs_init<<<vb, 256, 0, c->stream>>>(c->label.p, c->key.p, n);
  c->launches++;
  for (int round = 0; m > 0 && round < 1 << 20; ++round) {
    CK(cudaMemsetAsync(changed, 0, 4, c->stream));
    k_bfs_relax<<<std::max(eb, 1), 256, 0, c->stream>>>(c->edges.p, m, c->label.p, c->key.p, changed);
    c->launches++;
    CK(cudaMemcpyAsync(h_changed, changed, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (*h_changed == 0) break;
  }
  k_bfs_unpack<<<vb, 256, 0, c->stream>>>(c->key.p, c->generation.p, c->parent.p, n);
  c->launches++;
  CK(cudaGetLastError());
}

int swb200_d1_cluster(swb200_ctx *c, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent) {
  API_BEGIN(c)
  if (c->db_sharded) { g_err = "d1_cluster: not available on a sharded database (use the multi-GPU entry points)"; return SWB200_EUNSUPPORTED; }
  if (!c->have_network) { g_err = "d1_cluster: call swb200_d1_network first"; return SWB200_EINVAL; }
  const uint32_t n = c->n;
  c->tic();
  run_cluster(c);
  c->toc(3);
  c->clustered = true;
  {                                                           // rounds the relaxation took (stats[7]; persistent kernels only)
    uint32_t rounds = 0;
    if (c->cluster_kernel == 0 || c->cluster_kernel >= 3)
      CK(cudaMemcpyAsync(&rounds, reinterpret_cast<uint32_t *>(c->counters.p + 19), 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats[7] = rounds & 0x7FFFFFFFu;
  }
  if (swarm_of) CK(cudaMemcpyAsync(swarm_of, c->label.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (generation) CK(cudaMemcpyAsync(generation, c->generation.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (parent) CK(cudaMemcpyAsync(parent, c->parent.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END()
}

int swb200_d1_get_cluster(swb200_ctx *c, uint32_t first, uint32_t count, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent) {
  API_BEGIN(c)
  if (!c->clustered || static_cast<uint64_t>(first) + count > c->n) { g_err = "get_cluster: no clustering / bad range"; return SWB200_EINVAL; }
  if (swarm_of) CK(cudaMemcpyAsync(swarm_of, c->label.p + first, static_cast<size_t>(count) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (generation) CK(cudaMemcpyAsync(generation, c->generation.p + first, static_cast<size_t>(count) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (parent) CK(cudaMemcpyAsync(parent, c->parent.p + first, static_cast<size_t>(count) * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END()
}

uint64_t swb200_dist_buffer_bytes(uint32_t n_total, uint32_t world, uint32_t items_per_amplicon) {
  if (world == 0) return 0;
  const uint64_t S = (static_cast<uint64_t>(n_total) + world - 1) / world;
  // one sub-region per sender in both inboxes: links (8 B) and the message log of 16-byte records, which is sized
  // for kDistLogFactor records per link slot (a link is re-offered every time its source is lowered: ~2.6x measured)
  const uint64_t per_sender = std::max<uint64_t>(S * std::max<uint32_t>(items_per_amplicon, 1) / world, 1u << 14);
  return kDistCtlBytes + per_sender * world * (sizeof(uint2) + kDistLogFactor * sizeof(DistRec));
}

// rows of the result this rank owns: block-cyclic, blocks of kDistBlock ids
static uint32_t dist_own_rows(uint32_t n, uint32_t rank, uint32_t world) {
  const uint64_t nblocks = (static_cast<uint64_t>(n) + kDistBlock - 1) / kDistBlock;
  uint64_t rows = 0;
  for (uint64_t b = rank; b < nblocks; b += world) rows += std::min<uint64_t>(kDistBlock, n - b * kDistBlock);
  return static_cast<uint32_t>(rows);
}

uint32_t swb200_dist_row_count(uint32_t n_total, uint32_t rank, uint32_t world) {
  return (world == 0 || rank >= world) ? 0 : dist_own_rows(n_total, rank, world);
}

uint32_t swb200_dist_row_id(uint32_t rank, uint32_t world, uint32_t row) {
  return ((row / kDistBlock) * world + rank) * kDistBlock + (row % kDistBlock);
}

int swb200_dist_setup(swb200_ctx *c, uint32_t rank, uint32_t world, void *const *peer_buffers, uint64_t buffer_bytes) {
  API_BEGIN(c)
  if (world == 0 || world > kDistMaxWorld || rank >= world || !peer_buffers ||
      buffer_bytes < kDistCtlBytes + static_cast<uint64_t>(world) * 64 * (sizeof(uint2) + kDistLogFactor * sizeof(DistRec))) {
    g_err = "dist_setup: bad argument (1 <= world <= 16, one peer-visible buffer per rank)";
    return SWB200_EINVAL;
  }
  for (uint32_t r = 0; r < world; ++r) {
    if (!peer_buffers[r]) { g_err = "dist_setup: null peer buffer"; return SWB200_EINVAL; }
    c->dist_peer[r] = static_cast<unsigned char *>(peer_buffers[r]);
  }
  c->dist_rank = rank;
  c->dist_world = world;
  c->dist_buffer_bytes = buffer_bytes;
  c->idx_epoch = 0;
  c->dist_cap = (buffer_bytes - kDistCtlBytes) / (static_cast<uint64_t>(world) * (sizeof(uint2) + kDistLogFactor * sizeof(DistRec)));
  c->dist_cap &= ~1ull;                            // even: the 16-byte message records behind the link sub-regions stay aligned for any world size
  c->dist_calls = 0;
  c->dist_lcnt.alloc(2 * kDistMaxWorld);
  c->dist_links.alloc(c->dist_cap * world);
  CK(cudaMemsetAsync(c->dist_lcnt.p, 0, 2 * kDistMaxWorld * 8, c->stream));
  CK(cudaMemsetAsync(c->counters.p + 40, 0, 2 * 8, c->stream));
  CK(cudaMemsetAsync(c->dist_peer[rank], 0, kDistCtlBytes, c->stream));      // my own control block; the caller barriers before the first use
  CK(cudaStreamSynchronize(c->stream));
  API_END()
}

// Single-process multi-GPU: one context per device (or several per device: ranks sharing a GPU), driven by one host thread each.
// Allocates every rank's peer-visible buffer, opens peer access between the devices and hands all addresses to all contexts —
// what CUDA IPC / symmetric-memory handles do across processes is plain cudaMalloc + cudaDeviceEnablePeerAccess inside one.
int swb200_dist_setup_local(swb200_ctx *const *ctxs, uint32_t world, uint32_t n_total, uint32_t items_per_amplicon) {
  if (!ctxs || world == 0 || world > kDistMaxWorld) { g_err = "dist_setup_local: bad argument (1 <= world <= 16)"; return SWB200_EINVAL; }
  for (uint32_t r = 0; r < world; ++r) if (!ctxs[r]) { g_err = "dist_setup_local: null context"; return SWB200_EINVAL; }
  try {
    const uint64_t bytes = swb200_dist_buffer_bytes(n_total, world, items_per_amplicon);
    void *bufs[kDistMaxWorld] = {};
    for (uint32_t r = 0; r < world; ++r) {
      swb200_ctx *c = ctxs[r];
      CK(cudaSetDevice(c->device));
      for (uint32_t q = 0; q < world; ++q)
        if (ctxs[q]->device != c->device) {
          int can = 0;
          CK(cudaDeviceCanAccessPeer(&can, c->device, ctxs[q]->device));
          if (!can) { g_err = "dist_setup_local: the devices cannot access each other's memory (no NVLink / PCIe peer path)"; return SWB200_EUNSUPPORTED; }
          const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
          cudaGetLastError();
        }
      c->dist_own.alloc(bytes);
      CK(cudaMemset(c->dist_own.p, 0, bytes < (1u << 20) ? bytes : (1u << 20)));
      bufs[r] = c->dist_own.p;
      int sharing = 0;
      for (uint32_t q = 0; q < world; ++q) sharing += ctxs[q]->device == c->device ? 1 : 0;
      c->dist_grid_div = sharing;
    }
    for (uint32_t r = 0; r < world; ++r) {
      const int rc = swb200_dist_setup(ctxs[r], r, world, bufs, bytes);
      if (rc != SWB200_OK) return rc;
    }
  } catch (const CudaFail &f) {
    return f.code;
  }
  return SWB200_OK;
}

// every buffer swb200_d1_cluster_dist needs (no stream work)
static void dist_prepare(swb200_ctx *c) {
  const uint32_t n = c->n;
  const uint64_t nblocks = (static_cast<uint64_t>(n) + kDistBlock - 1) / kDistBlock;
  const uint32_t n_local = static_cast<uint32_t>((nblocks + c->dist_world - 1) / c->dist_world * kDistBlock);
  c->key.alloc(std::max<uint32_t>(n_local, 1)); c->label.alloc(n); c->generation.alloc(n); c->parent.alloc(n);
  c->cl_bits.alloc(static_cast<size_t>((n_local + 31) / 32) * 3);
  c->cl_ts.alloc(128);
  c->staging(4096);
  const uint32_t nblk = n_local / kDistBlock;
  c->bk_flag.alloc(static_cast<size_t>(nblk) * 3 + 16); c->bk_count.alloc(nblk + 1); c->bk_off.alloc(static_cast<size_t>(nblk) + 2);
  c->bk_links.alloc(std::max<uint64_t>(c->dist_cap * c->dist_world, 1));      // everything the link inboxes can hold
  c->bk_unit.alloc(c->bk_links.n / kBkUnit + 2); c->bk_act.alloc(c->bk_links.n / kBkUnit + 2);
}

// Allocate, ahead of time, every device buffer the multi-GPU step (d1_index with the index exchange, d1_network,
// d1_cluster_dist) will use for the loaded database.  A real multi-GPU job does not need this; it exists for jobs whose ranks
// SHARE one GPU (tests): cudaMalloc / cudaFree wait for the whole device, so a rank that allocated inside a step would block
// behind a peer's kernel that is itself spinning on a cross-rank barrier, waiting for that rank.
int swb200_d1_reserve(swb200_ctx *c) {
  API_BEGIN(c)
  if (c->n == 0 || c->dist_world == 0) { g_err = "d1_reserve: needs a database and swb200_dist_setup"; return SWB200_EINVAL; }
  c->jK = std::min<uint32_t>(64, c->min_len / 2);
  ts_prepare(c);
  dist_prepare(c);
  // CUDA loads a kernel lazily at its first launch, and that load synchronises the context: load them all now
  const void *kernels[] = {reinterpret_cast<const void *>(k_ts_wait), reinterpret_cast<const void *>(k_ts_route<true>),
                           reinterpret_cast<const void *>(k_ts_route<false>), reinterpret_cast<const void *>(k_ts_scatter_inbox<true>),
                           reinterpret_cast<const void *>(k_ts_scatter_inbox<false>), reinterpret_cast<const void *>(k_ts_scatter),
                           reinterpret_cast<const void *>(k_ts_join<true, true, 4>), reinterpret_cast<const void *>(k_ts_join<true, false, 4>),
                           reinterpret_cast<const void *>(k_ts_join<false, true, 4>), reinterpret_cast<const void *>(k_ts_join<false, false, 4>),
                           reinterpret_cast<const void *>(k_ts_join<true, true, 5>), reinterpret_cast<const void *>(k_ts_join<true, false, 5>),
                           reinterpret_cast<const void *>(k_ts_join<false, true, 5>), reinterpret_cast<const void *>(k_ts_join<false, false, 5>),
                           reinterpret_cast<const void *>(k_ts_join<true, true, 6>), reinterpret_cast<const void *>(k_ts_join<true, false, 6>),
                           reinterpret_cast<const void *>(k_ts_join<false, true, 6>), reinterpret_cast<const void *>(k_ts_join<false, false, 6>),
                           reinterpret_cast<const void *>(k_ts_big<true, true>), reinterpret_cast<const void *>(k_ts_big<true, false>),
                           reinterpret_cast<const void *>(k_ts_big<false, true>), reinterpret_cast<const void *>(k_ts_big<false, false>),
                           reinterpret_cast<const void *>(k_cluster_dist), reinterpret_cast<const void *>(k_cluster_bucket<true>),
                           reinterpret_cast<const void *>(k_cluster_bucket<false>),
                           reinterpret_cast<const void *>(k_ts_release), reinterpret_cast<const void *>(k_dist_rendezvous)};
  for (const void *k : kernels) {
    cudaFuncAttributes attr;
    CK(cudaFuncGetAttributes(&attr, k));
  }
  API_END()
}

int swb200_d1_cluster_dist(swb200_ctx *c, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent) {
  API_BEGIN(c)
  if (!c->have_network) { g_err = "d1_cluster_dist: call swb200_d1_network first"; return SWB200_EINVAL; }
  if (c->dist_world == 0) { g_err = "d1_cluster_dist: call swb200_dist_setup first"; return SWB200_EINVAL; }
  const uint32_t n = c->n;
  DistParams D{};
  D.rank = c->dist_rank; D.world = c->dist_world; D.n = n;
  const uint64_t nblocks = (static_cast<uint64_t>(n) + kDistBlock - 1) / kDistBlock;
  D.n_local = static_cast<uint32_t>((nblocks + D.world - 1) / D.world * kDistBlock);
  D.edges = c->edges.p; D.m_local = c->n_edges;
  dist_prepare(c);
  D.nwords = (D.n_local + 31) / 32;
  D.key = c->key.p; D.parent = c->parent.p; D.label = c->label.p; D.generation = c->generation.p; D.bits = c->cl_bits.p;
  D.my_links = c->dist_links.p; D.lcnt = c->dist_lcnt.p;
  for (uint32_t r = 0; r < D.world; ++r) D.peer[r] = c->dist_peer[r];
  D.cap_links = c->dist_cap;
  D.cap_upd = c->dist_cap * kDistLogFactor;
  D.lflags = reinterpret_cast<uint32_t *>(c->counters.p + 22);
  D.gbar = reinterpret_cast<unsigned int *>(c->counters.p + 40);
  if (const char *dbg = std::getenv("SWB200_DIST_DBG")) D.dbg = static_cast<uint32_t>(std::atoi(dbg));
  if (std::getenv("SWB200_CLUSTER_TS")) {
    c->cl_ts.alloc(128);
    D.ts = c->cl_ts.p;
  }
  const size_t dyn = static_cast<size_t>(kDistChunk) * sizeof(DistRec);
  BucketParams B{};
  B.nblk = D.n_local / kDistBlock;
  B.bcount = c->bk_count.p; B.boff = c->bk_off.p; B.blinks = c->bk_links.p; B.blinks_cap = c->bk_links.n;
  B.unit_blk = c->bk_unit.p; B.act_list = c->bk_act.p; B.unit_cap = c->bk_unit.n - 1; B.act_n = reinterpret_cast<uint32_t *>(c->counters.p + 46);
  // packed relaxation word (swarm | generation | parent, no parent pass: d1_bucket.cuh) while the ids of the job leave >= 10
  // generation bits; every rank derives the same answer from n, and a deeper swarm on ANY rank makes all of them go again unpacked
  bool pack = cluster_pack_bits(c, B.ib, B.gb) && c->dist_kernel != 1;
  uint32_t h[7] = {0, 0, 0, 0, 0, 0, 0};
  uint32_t overflow = 0;
  c->tic();
  for (;;) {
    D.epoch_base = (++c->dist_calls) << 24;
    if (D.ts) CK(cudaMemsetAsync(c->cl_ts.p, 0, 128 * 8, c->stream));
    B.D = D;
    const void *kern = c->dist_kernel == 1 ? reinterpret_cast<const void *>(k_cluster_dist)
                                           : (pack ? reinterpret_cast<const void *>(k_cluster_bucket<true>) : reinterpret_cast<const void *>(k_cluster_bucket<false>));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, dyn));
    const unsigned grid = std::max(1u, static_cast<unsigned>(c->sm_count * std::max(occ, 1)) / static_cast<unsigned>(c->dist_grid_div));
    void *args_d[] = {&D}, *args_b[] = {&B};
    if (c->dist_grid_div > 1) {                       // ranks sharing one GPU: cooperative kernels are never co-scheduled
      TsRouteParams R{};                              // ... and the ranks line up first (k_dist_rendezvous says why)
      R.rank = D.rank; R.world = D.world; R.epoch = c->dist_calls;
      for (uint32_t r = 0; r < R.world; ++r) R.peer[r] = c->dist_peer[r];
      R.err = D.lflags + 3;                           // a peer that never shows up is the barrier's error
      k_dist_rendezvous<<<1, 32, 0, c->stream>>>(R);
      c->launches++;
      if (c->dist_kernel == 1) k_cluster_dist<<<grid, 256, dyn, c->stream>>>(D);
      else if (pack) k_cluster_bucket<true><<<grid, 256, dyn, c->stream>>>(B);
      else k_cluster_bucket<false><<<grid, 256, dyn, c->stream>>>(B);
    } else {
      CK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), c->dist_kernel == 1 ? args_d : args_b, dyn, c->stream));
    }
    CK(cudaGetLastError());
    c->launches++;
    // read-backs go through PINNED memory: a copy into pageable memory blocks inside the driver until the kernel has finished,
    // and that kernel may be waiting for a peer whose host thread (ranks sharing one process) then cannot launch
    uint32_t *hp = static_cast<uint32_t *>(c->staging(4096));
    CK(cudaMemcpyAsync(hp, D.lflags, 7 * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(hp + 8, c->dist_peer[D.rank] + offsetof(DistCtl, overflow), 4, cudaMemcpyDeviceToHost, c->stream));
    c->toc(3);
    for (int i = 0; i < 7; ++i) h[i] = hp[i];
    overflow = hp[8];
    c->dist_rounds = h[4];
    if (D.ts) {
      unsigned long long t[128];
      CK(cudaMemcpy(t, c->cl_ts.p, sizeof t, cudaMemcpyDeviceToHost));
      std::string line = "[cluster_dist rank " + std::to_string(D.rank) + " rounds " + std::to_string(h[4]) +
                         "] us (" + std::string(c->dist_kernel == 1 ? "route, barrier, compact, then relax/barrier/apply per round, tail" : "init + route, count + scan, then relax + barrier / apply per round, tail") + "):";
      for (int i = 1; i < 128 && t[i]; ++i) line += " " + std::to_string(static_cast<long long>((t[i] - t[i - 1]) / 1000));
      std::fprintf(stderr, "%s\n", line.c_str());
    }
    if (h[3]) {
      CK(cudaMemsetAsync(c->counters.p + 40, 0, 2 * 8, c->stream));
      g_err = "d1_cluster_dist: a peer did not reach a barrier within 5 s";
      return SWB200_ECUDA;
    }
    // some rank met a generation beyond gb bits?  (the last barrier's vote, the same on every rank; a single rank reads its own flag)
    if (!pack || overflow || !(D.world > 1 ? h[5] : h[6])) break;
    pack = false;
    c->cluster_unpacked_reruns++;
  }
  if (overflow) {
    CK(cudaMemsetAsync(c->dist_peer[D.rank] + offsetof(DistCtl, overflow), 0, 4, c->stream));
    CK(cudaMemsetAsync(c->dist_lcnt.p, 0, 2 * kDistMaxWorld * 8, c->stream));
    g_err = "d1_cluster_dist: an inbox overflowed; set up larger peer buffers (swb200_dist_buffer_bytes with more items per amplicon)";
    return SWB200_ENOMEM;
  }
  c->clustered = true;
  // this rank's rows: block-cyclic (blocks of kDistBlock ids), packed in ascending id order
  const uint32_t rows = dist_own_rows(n, D.rank, D.world);
  const uint64_t full_blocks = rows / kDistBlock;
  auto fetch = [&](uint32_t *dst, const uint32_t *src) {
    if (!dst || rows == 0) return;
    if (full_blocks)
      CK(cudaMemcpy2DAsync(dst, kDistBlock * 4, src + static_cast<size_t>(D.rank) * kDistBlock, static_cast<size_t>(D.world) * kDistBlock * 4,
                           kDistBlock * 4, full_blocks, cudaMemcpyDeviceToHost, c->stream));
    const uint32_t tail = rows - static_cast<uint32_t>(full_blocks * kDistBlock);
    if (tail)
      CK(cudaMemcpyAsync(dst + full_blocks * kDistBlock, src + (full_blocks * D.world + D.rank) * kDistBlock, static_cast<size_t>(tail) * 4,
                         cudaMemcpyDeviceToHost, c->stream));
  };
  fetch(swarm_of, c->label.p);
  fetch(generation, c->generation.p);
  fetch(parent, c->parent.p);
  CK(cudaStreamSynchronize(c->stream));
  API_END()
}

int swb200_dn_cluster(swb200_ctx *c, uint32_t d, int no_cluster_breaking, const int64_t penalties[3], uint32_t *swarm_of,
                      uint32_t *generation, uint32_t *parent, uint32_t *pdiff) {
  API_BEGIN(c)
  if (c->n == 0 || !penalties || d < 2) { g_err = "dn_cluster: needs a database, penalties and d >= 2"; return SWB200_EINVAL; }
  if (c->db_sharded) { g_err = "dn_cluster: not available on a sharded database"; return SWB200_EUNSUPPORTED; }
  if (c->too_long) { g_err = "sequences longer than 5,000 nt are not supported by the clustering kernels"; return SWB200_EUNSUPPORTED; }
  const uint32_t n = c->n;
  const int64_t mis = penalties[0], go = penalties[1], ge = penalties[2];
  if (mis <= 0 || go < 0 || ge <= 0) { g_err = "dn_cluster: bad penalties"; return SWB200_EINVAL; }
  const int64_t B = static_cast<int64_t>(d) * std::max(mis, go + ge);
  const int64_t w64 = B >= go + ge ? (B - go) / ge : 0;
  if (B > (1 << 24)) { g_err = "dn_cluster: d x penalty exceeds 2^24"; return SWB200_EUNSUPPORTED; }
  const bool wide = w64 > 15;                     // beyond the register band of k_dn_align: k_dn_align_wide
  DnParams P{};
  P.words = c->words.p; P.len = c->len.p; P.abundance = c->abundance.p;
  P.n = n; P.stride = c->stride; P.d = d; P.ncb = no_cluster_breaking ? 1 : 0;
  P.mismatch = static_cast<int32_t>(mis); P.gapopen = static_cast<int32_t>(go); P.gapextend = static_cast<int32_t>(ge);
  P.bound = static_cast<int32_t>(B); P.w = static_cast<uint32_t>(std::min<int64_t>(w64, static_cast<int64_t>(c->max_len) + 1)); P.max_popc = 10 * d;
  P.max_len = c->max_len;
  c->qgrams.alloc(static_cast<size_t>(n) * 32);
  P.qgrams = c->qgrams.p;
  P.task_count = c->counters.p; P.edge_count = c->counters.p + 1; P.stats = c->counters.p + 2;
  c->tic();
  CK(cudaMemsetAsync(c->counters.p, 0, 8 * 8, c->stream));
  k_dn_qgrams<<<c->sm_count * 8, 256, 0, c->stream>>>(c->words.p, c->len.p, c->stride, n, c->qgrams.p);
  c->launches++;
  if (c->tasks.n == 0) c->tasks.alloc(std::max<size_t>(static_cast<size_t>(n) * 16, 1u << 20));
  unsigned long long ntasks = 0;
  const uint32_t Kp = std::min<uint32_t>(64, c->min_len / (d + 1));
  const bool use_join = Kp >= 8 && c->dn_filter != 1;
  DnJoinParams Q{};
  if (use_join) {
    const uint64_t slots = std::max<uint64_t>(64, (static_cast<uint64_t>(n) * (d + 1) * 5 / 2 + 3) / 4 * 4);
    c->jtab.alloc(slots);
    c->join_active = false;                       // the d=1 join table is overwritten
    c->indexed = false;
    Q.table = c->jtab.p; Q.n_buckets = slots / 4; Q.K = Kp;
    uint64_t seen = 1;
    while (seen < static_cast<uint64_t>(n) * 16) seen <<= 1;
    c->t2.alloc(seen);
    Q.seen = c->t2.p; Q.seen_mask = seen - 1;
    CK(cudaMemsetAsync(c->jtab.p, 0xFF, slots * 8, c->stream));
    k_dn_index_pieces<<<static_cast<unsigned>((static_cast<uint64_t>(n) * (d + 1) + 255) / 256), 256, 0, c->stream>>>(P, Q);
    c->launches++;
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    P.tasks = c->tasks.p; P.task_cap = c->tasks.n;
    CK(cudaMemsetAsync(c->counters.p, 0, 8, c->stream));
    if (use_join) {
      CK(cudaMemsetAsync(c->t2.p, 0xFF, (Q.seen_mask + 1) * 8, c->stream));
      k_dn_candidates_join<<<c->sm_count * 8, 256, 0, c->stream>>>(P, Q);
    } else {
      k_dn_filter<<<(n + 255) / 256, 256, 0, c->stream>>>(P, 0, (n + kDnTileQ - 1) / kDnTileQ);
    }
    c->launches++;
    CK(cudaMemcpyAsync(&ntasks, c->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (ntasks <= c->tasks.n) break;
    c->tasks.alloc(ntasks + ntasks / 8);
  }
  c->edges.alloc(std::max<uint64_t>(ntasks, 1));
  c->ediff.alloc(std::max<uint64_t>(ntasks, 1));
  P.edges = c->edges.p; P.ediff = c->ediff.p; P.edge_cap = c->edges.n;
  if (ntasks && wide) {
    // direction nibbles by absolute column, previous row's H and E by column: per thread max_len * (dir_words + 2) words
    P.dir_words = (std::max<uint32_t>(c->max_len, 1) + 7) / 8;
    const uint64_t per_thread = static_cast<uint64_t>(std::max<uint32_t>(c->max_len, 1)) * (P.dir_words + 2) * 4;
    uint64_t threads = std::min<uint64_t>(static_cast<uint64_t>(c->sm_count) * 16 * 128, (4ull << 30) / per_thread / 128 * 128);
    threads = std::max<uint64_t>(128, std::min<uint64_t>(threads, (ntasks + 127) / 128 * 128));
    c->dirs.alloc(threads * per_thread / 4);
    P.dirs = c->dirs.p;
    int32_t *hrow = reinterpret_cast<int32_t *>(c->dirs.p + threads * static_cast<uint64_t>(c->max_len) * P.dir_words);
    int32_t *erow = hrow + threads * static_cast<uint64_t>(c->max_len);
    k_dn_align_wide<<<static_cast<int>(threads / 128), 128, 0, c->stream>>>(P, hrow, erow, 0, ntasks);
    c->launches++;
    CK(cudaGetLastError());
  } else if (ntasks) {
    const uint32_t nb = 2 * P.w + 1;
    P.dir_words = (nb + 7) / 8;
    const uint64_t per_thread = static_cast<uint64_t>(std::max<uint32_t>(c->max_len, 1)) * P.dir_words * 4;
    uint64_t threads = static_cast<uint64_t>(c->sm_count) * 16 * 128;
    while (threads > 128 * static_cast<uint64_t>(c->sm_count) && threads * per_thread > (4ull << 30)) threads /= 2;
    threads = std::min<uint64_t>(threads, (ntasks + 127) / 128 * 128);
    c->dirs.alloc(threads * per_thread / 4);
    P.dirs = c->dirs.p;
    const int grid = static_cast<int>(threads / 128);
    if (P.w <= 4) k_dn_align<4><<<grid, 128, 0, c->stream>>>(P, 0, ntasks);
    else if (P.w <= 8) k_dn_align<8><<<grid, 128, 0, c->stream>>>(P, 0, ntasks);
    else k_dn_align<16><<<grid, 128, 0, c->stream>>>(P, 0, ntasks);
    c->launches++;
    CK(cudaGetLastError());
  }
  unsigned long long host[5];
  CK(cudaMemcpyAsync(host, c->counters.p, sizeof host, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->n_edges = host[1];
  c->dnstats[0] = host[2]; c->dnstats[1] = host[3]; c->dnstats[2] = host[4]; c->dnstats[3] = c->n_edges;
  c->have_network = true;
  run_cluster(c);
  c->pdiff.alloc(n);
  CK(cudaMemsetAsync(c->pdiff.p, 0, static_cast<size_t>(n) * 4, c->stream));
  if (c->n_edges) {
    k_dn_pdiff<<<c->sm_count * 4, 256, 0, c->stream>>>(c->edges.p, c->ediff.p, c->n_edges, c->parent.p, c->pdiff.p);
    c->launches++;
  }
  c->toc(6);
  c->clustered = true;
  if (swarm_of) CK(cudaMemcpyAsync(swarm_of, c->label.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (generation) CK(cudaMemcpyAsync(generation, c->generation.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (parent) CK(cudaMemcpyAsync(parent, c->parent.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (pdiff) CK(cudaMemcpyAsync(pdiff, c->pdiff.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END()
}

int swb200_d1_fastidious(swb200_ctx *c, uint64_t boundary, uint32_t *graft_cand, uint64_t *n_light, uint64_t *n_heavy) {
  API_BEGIN(c)
  if (!c->clustered) { g_err = "d1_fastidious: call swb200_d1_cluster first"; return SWB200_EINVAL; }
  if (c->db_sharded) { g_err = "d1_fastidious: not available on a sharded database"; return SWB200_EUNSUPPORTED; }
  const uint32_t n = c->n;
  const uint32_t Kj = std::min<uint32_t>(64, c->min_len / 3);
  if (c->fast_kernel != 1 && Kj >= 8) {
    // ---- pigeonhole join (d1_fastidious_join.cuh) ----
    c->mass.alloc(n); c->graft.alloc(n); c->is_light.alloc(n);
    uint32_t *counts_d = reinterpret_cast<uint32_t *>(c->counters.p + 10);
    const int vb = (n + 255) / 256;
    c->tic();
    CK(cudaMemsetAsync(c->mass.p, 0, static_cast<size_t>(n) * 8, c->stream));
    CK(cudaMemsetAsync(c->counters.p + 10, 0, 6 * 8, c->stream));
    k_fast_mass<<<vb, 256, 0, c->stream>>>(c->label.p, c->abundance.p, c->mass.p, n);
    k_fj_flags<<<vb, 256, 0, c->stream>>>(c->label.p, c->mass.p, boundary, n, c->is_light.p, c->graft.p, counts_d);
    c->launches += 2;
    uint32_t counts[2] = {0, 0};
    CK(cudaMemcpyAsync(counts, counts_d, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (n_light) *n_light = counts[0];
    if (n_heavy) *n_heavy = counts[1];
    if (counts[0] != 0 && counts[1] != 0) {
      JoinParams J{};
      J.words = c->words.p; J.len = c->len.p; J.n = n; J.stride = c->stride; J.K = Kj;
      J.is_light = c->is_light.p; J.graft_cand = c->graft.p; J.fstats = c->counters.p + 12;
      const uint64_t slots = std::max<uint64_t>(64, (static_cast<uint64_t>(counts[0]) * 3 * 5 / 2 + 3) / 4 * 4);
      c->t2.alloc(slots);
      J.table = c->t2.p; J.n_buckets = slots / 4;
      uint64_t bwords = 1024;
      while (bwords < static_cast<uint64_t>(counts[0]) * 3 / 4) bwords <<= 1;       // >= 16 bits per light piece
      c->fj_bloom.alloc(bwords);
      J.bloom = c->fj_bloom.p; J.bloom_words = bwords;
      CK(cudaMemsetAsync(c->fj_bloom.p, 0, bwords * 8, c->stream));
      CK(cudaMemsetAsync(c->t2.p, 0xFF, slots * 8, c->stream));
      k_fj_insert<<<vb, 256, 0, c->stream>>>(J);
      c->launches++;
      if (c->cands.n == 0) c->cands.alloc(std::max<size_t>(static_cast<size_t>(n) * 4, 1u << 20));
      J.cand_count = c->counters.p + 11;
      J.overflow = reinterpret_cast<uint32_t *>(c->counters.p + 48);
      // heavy amplicons in ascending id chunks (graft_cand[l] <= h prunes the later candidates of l); candidates and verification of
      // a chunk run back to back, the candidate count stays on the device — the host only looks at an overflow flag at the end
      uint32_t n_chunks = c->fast_chunks;
      for (int attempt = 0; attempt < 6; ++attempt) {
        CK(cudaMemsetAsync(c->counters.p + 48, 0, 8, c->stream));
        const uint32_t chunk = std::max<uint32_t>(1, (n + n_chunks - 1) / n_chunks);
        for (uint32_t a0 = 0; a0 < n; a0 += chunk) {
          const uint32_t a1 = static_cast<uint32_t>(std::min<uint64_t>(n, static_cast<uint64_t>(a0) + chunk));
          J.cands = c->cands.p; J.cand_cap = c->cands.n;
          CK(cudaMemsetAsync(c->counters.p + 11, 0, 8, c->stream));
          const uint64_t threads = static_cast<uint64_t>(a1 - a0) * 7;
          k_fj_candidates<<<static_cast<unsigned>(std::min<uint64_t>((threads + 255) / 256, static_cast<uint64_t>(c->sm_count) * 8)), 256, 0, c->stream>>>(J, a0, a1);
          k_fj_verify<<<c->sm_count * 8, 256, 0, c->stream>>>(J);
          c->launches += 2;
        }
        uint32_t ovf = 0;
        CK(cudaMemcpyAsync(&ovf, c->counters.p + 48, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (!ovf) break;
        // a chunk produced more candidates than the buffer holds: what was verified so far stands (every graft written is a real
        // one and the minimum is taken atomically); go again with smaller chunks and a larger buffer
        n_chunks *= 4;
        c->cands.alloc(c->cands.n * 2);
      }
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(c->fstats, c->counters.p + 12, 32, cudaMemcpyDeviceToHost, c->stream));
    c->toc(4);
    if (graft_cand) {
      CK(cudaMemcpyAsync(graft_cand, c->graft.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    return SWB200_OK;
  }
  c->mass.alloc(n); c->light_ids.alloc(n); c->heavy_ids.alloc(n); c->graft.alloc(n);
  FastParams F{};
  F.P = c->params();
  F.label = c->label.p; F.mass = c->mass.p; F.boundary = boundary;
  F.light_ids = c->light_ids.p; F.heavy_ids = c->heavy_ids.p;
  F.counts = reinterpret_cast<uint32_t *>(c->counters.p + 10);
  F.graft_cand = c->graft.p;
  F.fstats = c->counters.p + 12;
  const int vb = (n + 255) / 256;
  c->tic();
  CK(cudaMemsetAsync(c->mass.p, 0, static_cast<size_t>(n) * 8, c->stream));
  CK(cudaMemsetAsync(c->counters.p + 10, 0, 6 * 8, c->stream));
  k_fast_mass<<<vb, 256, 0, c->stream>>>(c->label.p, c->abundance.p, c->mass.p, n);
  k_fast_split<<<vb, 256, 0, c->stream>>>(F);
  c->launches += 2;
  uint32_t counts[2] = {0, 0};
  CK(cudaMemcpyAsync(counts, F.counts, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (n_light) *n_light = counts[0];
  if (n_heavy) *n_heavy = counts[1];
  if (counts[0] != 0 && counts[1] != 0) {        // otherwise nothing to graft (src/algod1.cc:1330-1334)
    // multimap sizing: <= 7L+4 variants per light amplicon, 2.5 slots per entry, 4 slots per bucket
    const uint64_t per_amp = 7ull * (static_cast<uint64_t>(c->stride) * 32) + 4;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    free_b += c->t2.n * 8;
    const uint64_t budget_slots = static_cast<uint64_t>(free_b * 0.85) / 8;
    uint64_t chunk = counts[0];
    if (static_cast<double>(chunk) * per_amp * 2.5 > static_cast<double>(budget_slots))
      chunk = std::max<uint64_t>(1, static_cast<uint64_t>(budget_slots / (per_amp * 2.5)));
    // exact variant count is <= 7*len+4 with the real lengths; use the mean bound when the set is one chunk
    const uint64_t slots = std::max<uint64_t>(64, static_cast<uint64_t>(static_cast<double>(chunk) * per_amp * 2.5) / 4 * 4);
    c->t2.alloc(slots);
    F.t2 = c->t2.p;
    F.n_buckets = slots / 4;
    const size_t zb = static_cast<size_t>(c->zlen) * 32;
    const size_t smem_l = zb + 8 * static_cast<size_t>(c->stride) * 8;
    const size_t smem_h = smem_l + 8 * sizeof(FastScratch);
    CK(cudaFuncSetAttribute(k_fast_light, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_l)));
    CK(cudaFuncSetAttribute(k_fast_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_h)));
    int occ_l = 1, occ_h = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l, k_fast_light, 256, smem_l));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_h, k_fast_heavy, 256, smem_h));
    for (uint64_t at = 0; at < counts[0]; at += chunk) {
      CK(cudaMemsetAsync(c->t2.p, 0xFF, slots * 8, c->stream));
      F.ids = c->light_ids.p + at;
      F.n_ids = static_cast<uint32_t>(std::min<uint64_t>(chunk, counts[0] - at));
      k_fast_light<<<c->sm_count * std::max(occ_l, 1), 256, smem_l, c->stream>>>(F);
      F.ids = c->heavy_ids.p;
      F.n_ids = counts[1];
      k_fast_heavy<<<c->sm_count * std::max(occ_h, 1), 256, smem_h, c->stream>>>(F);
      c->launches += 2;
    }
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(c->fstats, c->counters.p + 12, 32, cudaMemcpyDeviceToHost, c->stream));
  c->toc(4);
  if (graft_cand) {
    CK(cudaMemcpyAsync(graft_cand, c->graft.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  API_END()
}

// d = 0 (src/derep.cc:276-354): classes of identical sequences, kernels in d0_derep.cuh.  Device-timed as phase 7.
int swb200_d0_dereplicate(swb200_ctx *c, uint32_t *rep, uint64_t *mass, uint32_t *size, uint32_t *singletons, uint64_t *n_clusters) {
  API_BEGIN(c)
  if (c->db_sharded) { g_err = "d0_dereplicate: not available on a sharded database (use the multi-GPU entry points)"; return SWB200_EUNSUPPORTED; }
  if (c->n == 0) { g_err = "d0_dereplicate: no database loaded"; return SWB200_EINVAL; }
  if (c->db_pending) { g_err = "d0_dereplicate: swb200_load_db_shard must be followed by the row exchange and swb200_db_commit"; return SWB200_EINVAL; }
  const uint32_t n = c->n;
  const uint64_t slots = table_slots(n);                    // the reference's sizing rule (src/derep.cc:397)
  c->dr_table.alloc(slots); c->dr_slot.alloc(n); c->dr_rep.alloc(n); c->dr_mass.alloc(n); c->dr_size.alloc(n); c->dr_single.alloc(n);
  DerepParams D{};
  D.words = c->words.p; D.len = c->len.p; D.abundance = c->abundance.p; D.n = n; D.stride = c->stride;
  D.table = c->dr_table.p; D.slot_mask = slots - 1; D.slot_of = c->dr_slot.p; D.rep = c->dr_rep.p;
  D.mass = c->dr_mass.p; D.size = c->dr_size.p; D.singletons = c->dr_single.p;
  D.stats = c->counters.p + 24; D.count_steps = c->collect_stats;
  const unsigned blocks = (n + 255) / 256;
  c->tic();
  CK(cudaMemsetAsync(c->dr_table.p, 0xFF, slots * 8, c->stream));
  CK(cudaMemsetAsync(c->dr_mass.p, 0, static_cast<size_t>(n) * 8, c->stream));
  CK(cudaMemsetAsync(c->dr_size.p, 0, static_cast<size_t>(n) * 4, c->stream));
  CK(cudaMemsetAsync(c->dr_single.p, 0, static_cast<size_t>(n) * 4, c->stream));
  CK(cudaMemsetAsync(c->counters.p + 24, 0, 3 * 8, c->stream));
  k_derep_claim<<<blocks, 256, 0, c->stream>>>(D);
  k_derep_gather<<<blocks, 256, 0, c->stream>>>(D);
  CK(cudaGetLastError());
  c->launches += 2;
  c->toc(7);
  CK(cudaMemcpyAsync(c->drstats, c->counters.p + 24, 3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (rep) CK(cudaMemcpyAsync(rep, c->dr_rep.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (mass) CK(cudaMemcpyAsync(mass, c->dr_mass.p, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, c->stream));
  if (size) CK(cudaMemcpyAsync(size, c->dr_size.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (singletons) CK(cudaMemcpyAsync(singletons, c->dr_single.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (n_clusters) *n_clusters = c->drstats[0];
  API_END()
}

int swb200_stream(swb200_ctx *c, void **stream) {
  if (!c || !stream) { g_err = "null argument"; return SWB200_EINVAL; }
  *stream = reinterpret_cast<void *>(c->stream);
  return SWB200_OK;
}

double swb200_last_device_seconds(swb200_ctx *c) { return c ? c->last_s : 0.0; }
double swb200_phase_device_seconds(swb200_ctx *c, int phase) { return (c && phase >= 0 && phase < 8) ? c->phase_s[phase] : 0.0; }

int swb200_get_stats(swb200_ctx *c, uint64_t *out, int n) {
  if (!c || !out) { g_err = "null argument"; return SWB200_EINVAL; }
  c->stats[5] = c->launches;
  for (int i = 0; i < n && i < 8; ++i) out[i] = c->stats[i];
  for (int i = 8; i < n && i < 12; ++i) out[i] = c->fstats[i - 8];
  for (int i = 12; i < n && i < 16; ++i) out[i] = c->dnstats[i - 12];
  if (n > 16) out[16] = c->ts_overflow;
  if (n > 17) out[17] = c->ts_fallbacks;
  if (n > 18) out[18] = c->cluster_unpacked_reruns;
  return SWB200_OK;
}

int swb200_debug_variants(swb200_ctx *c, uint32_t seed, int mode, uint64_t *out_hash, uint32_t *out_code, uint32_t cap,
                          uint32_t *count, uint64_t *ztab, uint32_t ztab_cap, uint32_t *zlen) {
  API_BEGIN(c)
  if (c->n == 0 || seed >= c->n || !out_hash || !out_code || !count) { g_err = "debug_variants: bad argument"; return SWB200_EINVAL; }
  if (c->stride > 1024) { g_err = "debug_variants: sequence too long"; return SWB200_EUNSUPPORTED; }
  DevBuf<uint64_t> dh;
  DevBuf<uint32_t> dc, dn;
  dh.alloc(cap); dc.alloc(cap); dn.alloc(1);
  D1Params P = c->params();
  const size_t zbytes = static_cast<size_t>(c->zlen) * 32;
  if (mode == SWB200_ENUM_FULL) {
    CK(cudaFuncSetAttribute(k_d1_debug_variants<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(zbytes)));
    k_d1_debug_variants<0><<<1, 32, zbytes, c->stream>>>(P, seed, dh.p, dc.p, cap, dn.p);
  } else {
    CK(cudaFuncSetAttribute(k_d1_debug_variants<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(zbytes)));
    k_d1_debug_variants<1><<<1, 32, zbytes, c->stream>>>(P, seed, dh.p, dc.p, cap, dn.p);
  }
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(count, dn.p, 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const uint32_t m = std::min(*count, cap);
  CK(cudaMemcpy(out_hash, dh.p, static_cast<size_t>(m) * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_code, dc.p, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost));
  if (zlen) *zlen = c->zlen;
  if (ztab) std::memcpy(ztab, c->h_ztab.data(), std::min<size_t>(ztab_cap, c->h_ztab.size()) * 8);
  dh.release(); dc.release(); dn.release();
  API_END()
}

}  // extern "C"
