/* include/swarm_b200.h — C ABI of the B200 amplicon neighbour-search engine (libswarm_b200.so).
 *
 * The reference (torognes/swarm 3.1.6, /root/reference) has no plugin or FFI interface: its
 * clustering algorithms are C++ functions called once from main() after db_read() and they pull
 * their input through global accessors.  This header is the drop-in boundary a
 * maintainer would bind in their place; every entry point names the reference code it replaces.
 * Plain C: opaque context, caller-owned host buffers in, caller-owned host buffers out, int status,
 * thread-local error text.  The library never prints and never calls exit(): the host maps a
 * status to the reference's own message + exit code 1 (src/utils/fatal.h:27,38-46).
 *
 * Amplicon ids are indices into the database in the reference's order — abundance descending, then
 * strcmp(header) ascending (src/db.cc:392-406).  UINT32_MAX is "none" (`no_swarm`, src/algod1.cc:80).
 */
#ifndef SWARM_B200_H
#define SWARM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB200_NONE 0xFFFFFFFFu

enum {
  SWB200_OK = 0,
  SWB200_ECUDA = 1,        /* CUDA runtime error; text in swb200_last_error() */
  SWB200_EINVAL = 2,       /* bad argument / call order */
  SWB200_EDUPLICATE = 3,   /* identical sequences in the database (reference: fatal, src/algod1.cc:1141-1150) */
  SWB200_ENOMEM = 4,
  SWB200_EUNSUPPORTED = 5  /* e.g. sequence too long for the on-chip Zobrist table */
};

/* neighbour-enumeration strategy of swb200_d1_network (results are identical) */
enum {
  SWB200_ENUM_FULL = 0,  /* every seed probes all <= 7L+4 microvariants, exactly the reference's
                            enumeration (src/variants.cc:184-249) */
  SWB200_ENUM_HALF = 1,  /* each unordered neighbour pair is discovered once: deletions from the longer
                            sequence, substitutions from the side picked by the base tournament
                            0->1 0->2 1->2 1->3 2->3 3->0; both directed links are then derived from
                            the abundances.  ~3x fewer probes. */
  SWB200_ENUM_JOIN = 2   /* no microvariant enumeration: ed(u,v)=1 pairs share their first or last K
                            nucleotides, so two K-mer entries + two lookups per amplicon and an exact
                            packed-word comparison per candidate find the same links.  Needs the
                            shortest sequence to be >= 16 nt (falls back to HALF otherwise).  In this
                            mode duplicates are reported by swb200_d1_network.  Default. */
};

typedef struct swb200_ctx swb200_ctx;   /* owns the CUDA device, stream, and all device buffers */

/* Create a context on CUDA device `device` (one context per GPU / per process rank). */
int  swb200_create(swb200_ctx **out, int device);
void swb200_destroy(swb200_ctx *ctx);
const char *swb200_last_error(void);

/* Tunables (call before swb200_d1_index).  key: "enum_mode" (SWB200_ENUM_*), "bloom_bytes_per_slot"
 * (1,2,4,8: filter size = table slots * this; reference uses 1, src/algod1.cc:1127),
 * "collect_stats" (0/1), "net_kernel" (0 auto, 1 first-generation kernel, 2 lean HALF kernel),
 * "join_kernel" (JOIN mode: 0 auto = partitioned join over the tile store, d1_tilestore.cuh: one scatter pass writes
 * entry + packed row into fixed-capacity tile slots, a tile is joined from shared memory after one TMA bulk copy;
 * 1 = global hash multimap, d1_join.cuh, no length limit; 2 = r1's count / scan / scatter tile join, d1_tilejoin.cuh —
 * same links), "tile_cmax" (test hook: cap on the records a tile slot holds; the rest takes the overflow path),
 * "tile_rows" (0 default: a tile record is the 8-byte entry and the join gathers the packed rows it needs from the
 * database; 1: a record carries its packed row — the layout of the sharded multi-GPU job), "tile_cap" (tuning: records per
 * tile slot), "index_exchange" (after swb200_dist_setup: 2 = every rank hashes only its own rows and routes the records
 * to the tile owners over the peer buffers, 0 = every rank scans the whole replicated database and keeps its tiles' records,
 * 1 default = the exchange from 3 ranks on and always for a sharded database),
 * "skew_fallback" (1 default: when the overflow path would cost more than ~32 pair tests per amplicon — dense data,
 * huge groups sharing one K-mer — swb200_d1_network switches to the linear HALF enumeration, like the reference's
 * cost model; 0 = always sweep),
 * "fast_kernel" (0 auto, 1 microvariant multimap, 2 pigeonhole join — fastidious strategies, same result),
 * "fast_chunks" (tuning, default 8: the heavy amplicons are searched in this many ascending id chunks, so that a light amplicon
 * already grafted on an earlier heavy one prunes the later candidates),
 * "cluster_kernel" (0 = 5 fused label/generation relaxation over the link list in one persistent cooperative kernel;
 * 6 links bucketed by source block, inactive buckets skipped, d1_bucket.cuh (the multi-GPU kernel run on one GPU);
 * 4 frontier bitmap over 8-slot out-rows, d1_frontier.cuh;
 * 3 links counting-sorted by source; 2 one launch per round; 1 label propagation then BFS — same result),
 * "cluster_pack" (1 default: kernels 0 and 6 and the multi-GPU kernel relax ONE 64-bit word swarm | generation | parent, so
 * the parent is settled by the same atomicMin and no parent pass is needed; used while the ids leave >= 10 generation bits,
 * i.e. up to 2^27 amplicons, and redone with 32-bit generations if a swarm is deeper than that; 0 = key + parent pass),
 * "cluster_gen_bits" (test hook: cap on the generation bits of the packed word),
 * "dist_kernel" (multi-GPU clustering: 0 the bucketed kernel, 1 r1's k_cluster_dist), "dn_filter" (0 auto,
 * 1 all-pairs q-gram filter),
 * "shard_rank"/"shard_world" (this context's share of the network build, SURVEY.md §8e: in JOIN mode the
 * K-mer table is sharded by hash range — every rank scans all lookups but builds and walks only its own
 * bucket range; in the enumeration modes the seeds [rank*n/world, (rank+1)*n/world) are sharded). */
int  swb200_set_option(swb200_ctx *ctx, const char *key, int64_t value);

/* Replaces db_getsequence/len/abundance/count (src/db.h:35-53, src/db.cc:806-906) as the engine's
 * input: uploads the sorted, 2-bit packed database.  words: n*stride_words 64-bit words, amplicon i
 * at words[i*stride_words], 2 bits/nt LSB first (A0 C1 G2 T3, src/db.cc:100-114,561), zero padded;
 * len[i] in nt; abundance[i].  Host pointers; copied (pinned staging) to the device. */
int  swb200_load_db(swb200_ctx *ctx, const uint64_t *words, uint32_t stride_words,
                    const uint32_t *len, const uint64_t *abundance, uint32_t n);

/* The same database from fewer host bytes (the end-to-end path is PCIe-bound): lengths as 16-bit values (the engine
 * supports sequences up to 5 000 nt) and abundances as RUNS — the database is sorted by abundance (src/db.cc:392-406),
 * so amplicons run_start[r] .. run_start[r+1]-1 share run_abundance[r]; run_start has n_runs+1 entries, run_start[0] = 0,
 * run_start[n_runs] = n, strictly increasing.  Expanded on the device; everything downstream is unchanged.
 * 10 M x 150 bp: 420 MB instead of 520 MB. */
int  swb200_load_db_compact(swb200_ctx *ctx, const uint64_t *words, uint32_t stride_words, const uint16_t *len16,
                            const uint64_t *run_abundance, const uint32_t *run_start, uint32_t n_runs, uint32_t n);

/* Multi-GPU upload: every rank holds the whole database on its device, but pushes only its own
 * rows [first, first+count) over PCIe; the ranks then exchange rows device-to-device (NCCL all-gather over NVLink,
 * in place on the buffers swb200_db_device returns: equal shards of ceil(n_total / shard_world) rows, the buffers
 * are sized for that) and call swb200_db_commit.  Set "shard_world" before.  swb200_load_db_device adopts a database
 * that is already in device memory (device pointers, copied). */
int  swb200_load_db_shard(swb200_ctx *ctx, const uint64_t *words, uint32_t stride_words, const uint32_t *len,
                          const uint64_t *abundance, uint32_t n_total, uint32_t first, uint32_t count);
/* swb200_load_db_shard from fewer host bytes (the multi-GPU end-to-end path is bound by the host's aggregate PCIe rate): 16-bit
 * lengths for the rank's rows and NO abundance array — the abundance runs of the WHOLE database (as in swb200_load_db_compact)
 * are expanded on every device, so the row exchange only moves words and lengths (swb200_db_device: d_words, d_len). */
int  swb200_load_db_shard_compact(swb200_ctx *ctx, const uint64_t *words, uint32_t stride_words, const uint16_t *len16,
                                  uint32_t n_total, uint32_t first, uint32_t count, const uint64_t *run_abundance,
                                  const uint32_t *run_start, uint32_t n_runs);
/* SHARDED database (BASELINE configs[4], SURVEY.md §8e "hash-sharded"): this context keeps only the rows [first, first+count)
 * of the job's sorted database — no rank holds all packed sequences.  run_start (n_runs + 1 entries, as in
 * swb200_load_db_compact) describes the abundance runs of the WHOLE database: it is how a rank decides whether two amplicon ids
 * have equal abundances (src/algod1.cc:580-583) without the other ranks' rows.  Requires swb200_dist_setup, option
 * "tile_rows" = 1 (records carry their packed row to the rank that owns their tile) and, when the ranks' shards differ in
 * their shortest / longest sequence, the options "job_min_len" / "job_max_len" (set before this call; every rank must derive
 * the same piece length).  Then swb200_d1_index (the index exchange runs inside it, over the peer buffers),
 * swb200_d1_network and swb200_d1_cluster_dist; the single-GPU entry points answer SWB200_EUNSUPPORTED. */
int  swb200_load_db_rows(swb200_ctx *ctx, const uint64_t *words, uint32_t stride_words, const uint32_t *len,
                         const uint64_t *abundance, uint32_t n_total, uint32_t first, uint32_t count,
                         const uint32_t *run_start, uint32_t n_runs);
int  swb200_db_device(swb200_ctx *ctx, void **d_words, void **d_len, void **d_abundance);
int  swb200_db_commit(swb200_ctx *ctx);
int  swb200_load_db_device(swb200_ctx *ctx, const void *d_words, uint32_t stride_words, const void *d_len,
                           const void *d_abundance, uint32_t n);

/* Replaces the "Hashing sequences" phase: zobrist_hash of every amplicon (src/db.cc:761,
 * src/zobrist.cc:134-184), hash_alloc + hash_insert + bloom_set (src/algod1.cc:1118-1139,188-208),
 * duplicate detection (:174-185).  Returns SWB200_EDUPLICATE when two amplicons are identical. */
int  swb200_d1_index(swb200_ctx *ctx);

/* Replaces the "Building network" phase: network_thread -> check_variants -> generate_variants +
 * find_variant_matches + check_variant (src/algod1.cc:558-670, src/variants.cc:118-249,
 * src/bloompat.cc:68-71, src/hashtable.cc:47-87).  A directed link u->v exists iff v is a
 * microvariant of u and (no_cluster_breaking or abundance(u) >= abundance(v)) (:580-583).
 * n_links: number of directed links found by this context's shard. */
int  swb200_d1_network(swb200_ctx *ctx, int no_cluster_breaking, uint64_t *n_links);

/* The network as CSR = `ampinfo[i].link_start/link_count` + `network_v` (src/algod1.cc:94-95,162),
 * rows sorted ascending like the -j writer (:769-772).  row_ptr: n+1 entries; col: n_links. */
int  swb200_d1_get_network(swb200_ctx *ctx, uint64_t *row_ptr, uint32_t *col);

/* Multi-GPU exchange step: raw link list of this shard as (src,dst) pairs, and
 * replacement of the context's link list by the gathered one.  pairs: 2*n_links uint32. */
int  swb200_d1_export_links(swb200_ctx *ctx, uint32_t *pairs);
int  swb200_d1_import_links(swb200_ctx *ctx, const uint32_t *pairs, uint64_t n_links);
/* device-resident variants for NCCL all-gatherv (pointers are CUDA device addresses) */
int  swb200_d1_links_device(swb200_ctx *ctx, void **d_pairs, uint64_t *n_links);
int  swb200_d1_import_links_device(swb200_ctx *ctx, const void *d_pairs, uint64_t n_links);

/* Replaces the "Clustering" phase (greedy BFS, src/algod1.cc:1185-1280, process_seed :673-718):
 * swarm_of[i] = amplicon id of the seed of i's swarm; generation[i] = BFS depth; parent[i] =
 * the amplicon that claimed i (SWB200_NONE for seeds).  Any output pointer may be NULL. */
int  swb200_d1_cluster(swb200_ctx *ctx, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent);

/* The rows [first, first+count) of the last clustering (a rank's own share of the result). */
int  swb200_d1_get_cluster(swb200_ctx *ctx, uint32_t first, uint32_t count, uint32_t *swarm_of, uint32_t *generation,
                           uint32_t *parent);

/* Multi-GPU clustering of ONE job.  Each rank calls swb200_d1_network on its share of the join
 * (shard_rank / shard_world), then swb200_d1_cluster_dist on all ranks together: one persistent kernel per GPU
 * routes the links to the owners of their sources, relaxes them in rounds and exchanges the cross-rank updates
 * itself, over NVLink, by writing into the peers' inboxes (no link gather, no host round trip per round).
 * Ownership is block-cyclic: rank r owns the ids of the blocks b = r, r+world, ... of 4096 consecutive ids; its result
 * arrays hold those rows packed in ascending id order (swb200_dist_row_count rows; row i is amplicon
 * swb200_dist_row_id(rank, world, i)).  The inboxes are peer-visible buffers the CALLER provides (CUDA IPC / torch
 * symmetric memory): peer_buffers[r] = address, in this process, of rank r's buffer; every buffer has buffer_bytes >=
 * swb200_dist_buffer_bytes(n, world, items_per_amplicon) (4 is ample unless the network is very dense; an overflow is
 * reported as SWB200_ENOMEM).  Call swb200_dist_setup on every rank, then a barrier, before the first clustering. */
uint64_t swb200_dist_buffer_bytes(uint32_t n_total, uint32_t world, uint32_t items_per_amplicon);
uint32_t swb200_dist_row_count(uint32_t n_total, uint32_t rank, uint32_t world);
uint32_t swb200_dist_row_id(uint32_t rank, uint32_t world, uint32_t row);
int  swb200_dist_setup(swb200_ctx *ctx, uint32_t rank, uint32_t world, void *const *peer_buffers, uint64_t buffer_bytes);
/* The same for ONE process that drives all GPUs (the command line, SWARM_B200_DEVICES): ctxs[r] = rank r's context (one per
 * device, or several on one device — they then share its SMs); the peer buffers are allocated here and peer access between
 * the devices is enabled.  Each rank is then driven by its own host thread. */
int  swb200_dist_setup_local(swb200_ctx *const *ctxs, uint32_t world, uint32_t n_total, uint32_t items_per_amplicon);
int  swb200_d1_cluster_dist(swb200_ctx *ctx, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent);
/* Optional: allocate now every device buffer the multi-GPU step will use for the loaded database.  Only needed when several
 * ranks SHARE one GPU (tests): cudaMalloc / cudaFree wait for the whole device, so allocating inside a step would block behind
 * a peer's kernel that is spinning on a cross-rank barrier. */
int  swb200_d1_reserve(swb200_ctx *ctx);

/* Replaces the fastidious passes (mark_light_thread / check_heavy_thread, src/algod1.cc:374-552):
 * for every amplicon l of a light swarm (mass < boundary), graft_cand[l] = the smallest amplicon id h
 * belonging to a heavy swarm such that h and l share a microvariant, else SWB200_NONE (:244-258).
 * The attach order (:274-336) is applied by the host from this array.  Requires swb200_d1_cluster.
 * Returns via n_light / n_heavy the amplicon counts; graft_cand is all NONE when either is zero. */
int  swb200_d1_fastidious(swb200_ctx *ctx, uint64_t boundary, uint32_t *graft_cand,
                          uint64_t *n_light_amplicons, uint64_t *n_heavy_amplicons);

/* d>1: replaces algo_run's clustering (src/algo.cc:329-708): db_qgrams_init/findqgrams
 * (src/qgram.cc:68-96), the q-gram prefilter (qgram_diff, src/qgram.cc:247-252), search_do -> search8 /
 * search16 + backtrack (src/scan.cc:221-256, src/search16.cc:207-230, src/utils/backtrack.h:51-138) and
 * the greedy loop (src/algo.cc:384-602).  penalties = converted mismatch, gap-open, gap-extend costs
 * (src/swarm.cc:466-483; defaults 18, 24, 13).  Outputs as swb200_d1_cluster plus pdiff[i] = number of
 * differences between i and its parent (the -i file's third column).  Needs only swb200_load_db.
 * Duplicate sequences must be rejected by the host at d>1 (src/db.cc:763-796). */
int  swb200_dn_cluster(swb200_ctx *ctx, uint32_t d, int no_cluster_breaking, const int64_t penalties[3],
                       uint32_t *swarm_of, uint32_t *generation, uint32_t *parent, uint32_t *pdiff);

/* d=0, dereplication: replaces `dereplicating` (src/derep.cc:276-354: zobrist_hash + open-addressing table of
 * clusters + exact sequence comparison, chains in index order).  rep[i] = the smallest id with i's sequence (the
 * reference's seqno_first of i's cluster); mass / size / singletons are the cluster sums (src/derep.cc:322-344),
 * stored at the representative's index and 0 elsewhere; any output pointer may be NULL.  The host orders clusters by
 * mass descending, then representative (sort_seeds, :74-98).  Needs only swb200_load_db.  Device time: phase 7. */
int  swb200_d0_dereplicate(swb200_ctx *ctx, uint32_t *rep, uint64_t *mass, uint32_t *size, uint32_t *singletons,
                           uint64_t *n_clusters);

/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on: lets the caller bracket calls
 * with its own CUDA events, or order collectives (NCCL) after the engine's work without a host round trip. */
int  swb200_stream(swb200_ctx *ctx, void **stream);

/* Device time (CUDA events on the engine's stream) of the last call, and accumulated per phase.
 * phase: 0 load_db(H2D) 1 index 2 network 3 cluster 4 fastidious 6 d>1 (all of swb200_dn_cluster)
 *        7 d=0 (swb200_d0_dereplicate) */
double swb200_last_device_seconds(swb200_ctx *ctx);
double swb200_phase_device_seconds(swb200_ctx *ctx, int phase);

/* Counters of the last network build (collect_stats=1): [0] variants probed, [1] filter passes,
 * [2] table slots visited, [3] exact comparisons, [4] links, [5] kernel launches since create, [6] packed sequences gathered into shared memory (tile join), [7] rounds of the last swb200_d1_cluster; [8..11] fastidious: light variants stored, heavy variants probed, tag matches, verified matches;
 * [12..15] d>1: q-gram comparisons, alignments, alignments pruned early, accepted links; [16] tile-store records that
 * overflowed their tile slot in the last index, [17] times the network fell back to the enumeration (skew_fallback),
 * [18] times the clustering was redone with 32-bit generations (a swarm deeper than the packed word holds, "cluster_pack"). */
int  swb200_get_stats(swb200_ctx *ctx, uint64_t *out, int n);

/* Test hook: enumerate the microvariants of amplicon `seed` on the device exactly as the network
 * kernel does (mode = SWB200_ENUM_*); out_hash/out_code hold up to cap entries, code =
 * type<<30 | base<<28 | pos (type 0 sub, 1 del, 2 ins).  Also returns the engine's Zobrist table
 * (zlen*4 values, position-major) when ztab != NULL. */
int  swb200_debug_variants(swb200_ctx *ctx, uint32_t seed, int mode, uint64_t *out_hash,
                           uint32_t *out_code, uint32_t cap, uint32_t *count,
                           uint64_t *ztab, uint32_t ztab_cap, uint32_t *zlen);

#ifdef __cplusplus
}
#endif
#endif
