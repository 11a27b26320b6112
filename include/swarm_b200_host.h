/* include/swarm_b200_host.h — C ABI of the host layer (libswarm_b200_host.so): the FASTA database
 * and the output writers, i.e. the host mirror of the reference's src/db.cc and of the writers in
 * src/algod1.cc:755-1095.  No CUDA here; the arrays it exposes are exactly the arguments of
 * swb200_load_db() (include/swarm_b200.h).
 */
#ifndef SWARM_B200_HOST_H
#define SWARM_B200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct swbh_db swbh_db;

const char *swbh_last_error(void);

/* db_read (src/db.cc:432-803): parse, 2-bit pack, abundance annotations, uniqueness checks, sort.
 * usearch_abundance = -z, append_abundance = -a (0 = off), check_dup_sequences = d>1 behaviour.
 * Returns 0, or 1 with the reference's error text (without the "\nError: " prefix) in swbh_last_error(). */
int  swbh_db_read_fasta(const char *path, int usearch_abundance, int64_t append_abundance,
                        int check_dup_sequences, swbh_db **out);
int  swbh_db_parse(const char *text, uint64_t size, int usearch_abundance, int64_t append_abundance,
                   int check_dup_sequences, swbh_db **out);
void swbh_db_free(swbh_db *db);
/* workers of the FASTA ingest (0 = hardware concurrency, at most 32; 1 = serial).  Results do not depend on it. */
void swbh_set_threads(int threads);
/* The output writers below format large results with the same workers (swarms / rows cut into ranges of equal work, texts
 * concatenated in order: identical bytes for any worker count).  Test hook: the amount of work from which they do so. */
void swbh_set_writer_grain(uint64_t weight);

uint32_t swbh_db_count(const swbh_db *db);            /* db_getsequencecount   src/db.cc:806-809 */
uint32_t swbh_db_longest(const swbh_db *db);          /* db_getlongestsequence src/db.cc:812-815 */
uint32_t swbh_db_stride_words(const swbh_db *db);
uint64_t swbh_db_nucleotides(const swbh_db *db);
const uint64_t *swbh_db_words(const swbh_db *db);     /* db_getsequence  (fixed stride)           */
const uint32_t *swbh_db_lengths(const swbh_db *db);   /* db_getsequencelen                        */
const uint64_t *swbh_db_abundances(const swbh_db *db);/* db_getabundance                          */
const char *swbh_db_header(const swbh_db *db, uint32_t i);  /* db_getheader                       */
/* the arguments of swb200_load_db_compact(): 16-bit lengths + abundance runs (built on first use, owned by db).
 * Returns n_runs, or 0 when a sequence is longer than 65 535 nt (use swb200_load_db then). */
uint32_t swbh_db_compact(swbh_db *db, const uint16_t **len16, const uint64_t **run_abundance, const uint32_t **run_start);

/* d=1 result assembly + writers (src/algod1.cc:791-815 `-o`, :1043-1062 `-s`, :990-1040 `-i`,
 * :755-788 `-j`, :937-987 `-w`, :818-849 `-r`).  Inputs are the engine's outputs: swarm_of /
 * generation / parent per amplicon, and graft_cand (may be NULL = no fastidious).  Each writer
 * appends to a malloc'ed buffer returned through out/out_len (free with swbh_free). */
typedef struct swbh_result swbh_result;
int  swbh_d1_assemble(const swbh_db *db, const uint32_t *swarm_of, const uint32_t *generation,
                      const uint32_t *parent, const uint32_t *graft_cand, uint64_t boundary,
                      swbh_result **out);
void swbh_result_free(swbh_result *r);
uint64_t swbh_result_swarms(const swbh_result *r);    /* swarm count after grafting */
uint32_t swbh_result_largest(const swbh_result *r);
uint32_t swbh_result_maxgen(const swbh_result *r);
uint64_t swbh_result_grafts(const swbh_result *r);
int  swbh_write_swarms(const swbh_db *db, const swbh_result *r, int mothur, int64_t differences,
                       int usearch_abundance, int64_t append_abundance, char **out, uint64_t *out_len);
int  swbh_write_stats(const swbh_db *db, const swbh_result *r, int usearch_abundance, char **out, uint64_t *out_len);
int  swbh_write_structure(const swbh_db *db, const swbh_result *r, int usearch_abundance, char **out, uint64_t *out_len);
int  swbh_write_seeds(const swbh_db *db, const swbh_result *r, int usearch_abundance, char **out, uint64_t *out_len);
int  swbh_write_network(const swbh_db *db, const uint64_t *row_ptr, const uint32_t *col,
                        int usearch_abundance, int64_t append_abundance, char **out, uint64_t *out_len);
/* d>1 result assembly + writers (src/algo.cc:259-325 `-o/-r`, :608-674 `-s`, :470-484,:573-586 `-i`).
 * pdiff[i] = differences between i and its parent; radius is accumulated here.  The swarm lists use
 * the same order as d=1 (seed, then generations, ids ascending: src/algo.cc:205-256), so
 * swbh_write_swarms() serves both. */
int  swbh_dn_assemble(const swbh_db *db, const uint32_t *swarm_of, const uint32_t *generation,
                      const uint32_t *parent, const uint32_t *pdiff, swbh_result **out);
int  swbh_dn_write_stats(const swbh_db *db, const swbh_result *r, int usearch_abundance, char **out, uint64_t *out_len);
int  swbh_dn_write_structure(const swbh_db *db, const swbh_result *r, int usearch_abundance, char **out, uint64_t *out_len);
/* -u, the UCLUST-like records (src/algod1.cc:851-934 at d=1, src/algo.cc:608-661 at d>1): C and S per swarm, one H per
 * other member with the reference's scalar global alignment against the seed (src/nw.cc:40-252, CIGAR
 * src/utils/cigar.cc:30-60).  penalties = swbh_scoring()'s output; `threads` workers align (output independent of it). */
int  swbh_write_uclust(const swbh_db *db, const swbh_result *r, int64_t differences, const int64_t penalties[3],
                       int usearch_abundance, int64_t append_abundance, int threads, char **out, uint64_t *out_len);
/* d=0 (dereplication) result assembly + writers, src/derep.cc: clusters ordered by mass descending then seed index
 * (sort_seeds :74-98), members in index order (:322-326); -o/-r :206-273, -w :190-203, -u :145-187, -i :121-142,
 * -s :103-118.  Inputs are swb200_d0_dereplicate()'s outputs. */
typedef struct swbh_derep swbh_derep;
int  swbh_d0_assemble(const swbh_db *db, const uint32_t *rep, const uint64_t *mass, const uint32_t *size,
                      const uint32_t *singletons, swbh_derep **out);
void swbh_derep_free(swbh_derep *r);
uint64_t swbh_derep_clusters(const swbh_derep *r);
uint32_t swbh_derep_largest(const swbh_derep *r);
uint64_t swbh_derep_heaviest(const swbh_derep *r);    /* "Heaviest swarm" of the log, src/derep.cc:417 */
int  swbh_d0_write_swarms(const swbh_db *db, const swbh_derep *r, int mothur, int usearch_abundance, int64_t append_abundance,
                          char **out, uint64_t *out_len);
int  swbh_d0_write_seeds(const swbh_db *db, const swbh_derep *r, int usearch_abundance, char **out, uint64_t *out_len);
int  swbh_d0_write_uclust(const swbh_db *db, const swbh_derep *r, int usearch_abundance, int64_t append_abundance,
                          char **out, uint64_t *out_len);
int  swbh_d0_write_structure(const swbh_db *db, const swbh_derep *r, int usearch_abundance, char **out, uint64_t *out_len);
int  swbh_d0_write_stats(const swbh_db *db, const swbh_derep *r, int usearch_abundance, char **out, uint64_t *out_len);
/* first band half-width of the -u aligner (default 8; it doubles until the banded result provably equals the full
 * matrix's; 0 = always the full matrix).  The records do not depend on it. */
void swbh_uclust_band(int first_half_width);
/* alignment scoring conversion (src/swarm.cc:466-483): penalties[3] = mismatch, gap open, gap extend */
void swbh_scoring(int64_t match_reward, int64_t mismatch_penalty, int64_t gap_open, int64_t gap_extend, int64_t penalties[3]);

void swbh_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
