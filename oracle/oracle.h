/* oracle/oracle.h — CPU restatement of swarm's neighbour-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under swarm_b200/ (the product) may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and only as the checker / reported baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against outputs
 * of the unmodified reference binary (oracle/_ref/swarm, built by oracle/Makefile from
 * /root/reference) committed under tests/golden/ together with tests/golden/make_golden.py.
 *
 * Each function cites the reference file:line (relative to /root/reference) it restates.  The code
 * is written from scratch in plain C11 over an SoA database (fixed: 2 bits/nt, LSB first, A0 C1 G2
 * T3 — src/db.cc:100-114,561; sequence i = words[off[i] .. off[i] + ceil(len[i]/32))).
 */
#ifndef SWARM_ORACLE_H
#define SWARM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NONE 0xFFFFFFFFu            /* `no_swarm` = UINT_MAX, src/algod1.cc:80 */

typedef struct {
  uint32_t n;                /* amplicons, already sorted (abundance desc, header asc): src/db.cc:392-406 */
  uint32_t longest;          /* longest sequence, nt */
  const uint64_t *words;     /* packed sequences */
  const uint64_t *off;       /* n+1 word offsets */
  const uint32_t *len;       /* n lengths (nt) */
  const uint64_t *abundance; /* n abundances */
} orc_db;

/* variant record, same meaning as `var_s` (src/variants.h:31-38); type 0 sub, 1 del, 2 ins */
typedef struct { uint64_t hash; uint32_t pos; uint8_t type; uint8_t base; uint16_t pad; } orc_var;

/* --- Zobrist hashing: src/zobrist.cc:49-80,127-131,134-240; RNG src/utils/pseudo_rng.h:30-31 --- */
void     orc_zobrist_init(uint32_t zobrist_len);     /* re-seeds mt19937_64(1), fills 4*len values */
void     orc_zobrist_exit(void);
uint64_t orc_zobrist_value(uint32_t pos, uint32_t base);
uint64_t orc_zobrist_hash(const uint64_t *seq, uint32_t len);
uint64_t orc_mt19937_64_next(void);                  /* exposed for the known-answer test */
void     orc_mt19937_64_seed(uint64_t seed);         /* std::mt19937_64 default seed = 5489 */

/* --- microvariants: src/variants.cc:184-249 (enumeration), :118-165 (verification), :78-115 --- */
uint32_t orc_generate_variants(const uint64_t *seq, uint32_t len, uint64_t hash, orc_var *out);
int      orc_check_variant(const uint64_t *seed, uint32_t seedlen, const orc_var *v,
                           const uint64_t *amp, uint32_t amplen);
void     orc_generate_variant_sequence(const uint64_t *seed, uint32_t seedlen, const orc_var *v,
                                       uint64_t *out, uint32_t *outlen);

/* --- hash-table size: src/utils/hashtable_size.cc:29-42 --- */
uint64_t orc_hashtable_size(uint64_t n);

/* --- d=1 network: src/algod1.cc:1118-1171 (index + network), :558-627.
 * Returns 0, or 1 if two amplicons have identical sequences (the reference aborts: :1141-1150).
 * link_start/link_count: n entries each; *network is malloc'ed (free with orc_free), rows hold
 * neighbour ids in variant-enumeration order exactly like `network_v`. */
int orc_d1_network(const orc_db *db, int no_cluster_breaking,
                   uint32_t *link_start, uint32_t *link_count, uint32_t **network, uint64_t *n_edges,
                   uint64_t *stats /* NULL or [4]: variants, bloom passes, slots visited, exact compares */);

/* --- d=1 greedy clustering: src/algod1.cc:1185-1280, :673-718, :746-752.
 * swarmid[i] = running swarm number; generation, parent (ORC_NONE for seeds), next = linked list.
 * Per-swarm arrays (capacity n): seed, last, size, singletons, maxgen, mass, sumlen.  Returns #swarms. */
uint32_t orc_d1_cluster(const orc_db *db, const uint32_t *link_start, const uint32_t *link_count,
                        const uint32_t *network,
                        uint32_t *swarmid, uint32_t *generation, uint32_t *parent, uint32_t *next,
                        uint32_t *sw_seed, uint32_t *sw_last, uint32_t *sw_size, uint32_t *sw_singletons,
                        uint32_t *sw_maxgen, uint64_t *sw_mass, uint64_t *sw_sumlen);

/* --- fastidious: src/algod1.cc:1291-1475 with :214-336 (attach), :339-552 (light / heavy passes),
 * src/bloomflex.cc:43-115.  Inputs are orc_d1_cluster's outputs (modified in place the way the
 * reference does: `next`, sw_last/size/singletons/mass/sumlen of heavy swarms, sw_attached).
 * graft_cand[n]: final value per amplicon (ORC_NONE if none / cleared, :320-324).
 * Returns number of grafts, or -1 when there are only light or only heavy swarms (:1330-1334). */
int64_t orc_d1_fastidious(const orc_db *db, uint64_t boundary, uint32_t bloom_bits, uint32_t nswarms,
                          const uint32_t *swarmid, uint32_t *next,
                          uint32_t *sw_seed, uint32_t *sw_last, uint32_t *sw_size, uint32_t *sw_singletons,
                          uint64_t *sw_mass, uint64_t *sw_sumlen, uint8_t *sw_attached,
                          uint32_t *graft_cand, uint32_t *graft_raw /* NULL or n: min heavy id before the attach loop */,
                          uint64_t *stats /* NULL or [4] */);

/* --- d>1 (oracle_dn.c): q-grams src/qgram.cc:68-96,247-252; scoring src/swarm.cc:466-483; aligner
 * src/nw.cc:40-191; greedy loop src/algo.cc:384-602 --- */
void     orc_findqgrams(const uint64_t *seq, uint32_t len, uint8_t *vec /* 128 bytes */);
uint64_t orc_qgram_diff(const uint8_t *a, const uint8_t *b);
void     orc_scoring(int64_t match, int64_t mismatch, int64_t gapopen, int64_t gapextend, int64_t out[3]);
uint64_t orc_nw_diffs(const uint64_t *dseq, uint32_t dlen, const uint64_t *qseq, uint32_t qlen,
                      int64_t mismatch, int64_t gapopen, int64_t gapextend, uint64_t *alnlen);
uint32_t orc_dn_cluster(const orc_db *db, uint32_t d, int no_cluster_breaking, const int64_t pen[3],
                        uint32_t *order, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent,
                        uint32_t *pdiff, uint32_t *radius, uint64_t *stats /* NULL or [3] */);

/* --- d=0 (oracle_d0.c): src/derep.cc:276-354 + sort_seeds :74-98.  rep[n] = seed of a's cluster, next[n] = chain in
 * index order (0 ends it), per cluster in output order: seeds, mass, size, singletons (capacity n).  Returns #clusters. */
uint32_t orc_d0_dereplicate(const orc_db *db, uint32_t *rep, uint32_t *next, uint32_t *seeds, uint64_t *mass,
                            uint32_t *size, uint32_t *singletons);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
