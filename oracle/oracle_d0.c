/* oracle/oracle_d0.c — CPU restatement of swarm's d = 0 dereplication (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 * Follows /root/reference src/derep.cc:276-354 (`dereplicating`) step by step: amplicons are visited in index
 * order; the bucket is found by linear probing from hash & mask in a table of compute_hashtable_size(n) buckets
 * (src/derep.cc:397, src/utils/hashtable_size.cc:29-42) and a bucket matches when hash, length and packed words are
 * equal (:303-318); the first amplicon of a cluster is its seed (seqno_first), later ones are appended to the chain
 * `nextseqtab` (:322-344).  Then the clusters are ordered by mass descending, seed ascending (sort_seeds, :74-98).
 * Parity: pinned against the reference binary's -d 0 outputs (tests/golden/NAME.d0.EXT) through the host writers.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t hash, mass; uint32_t first, last, size, singletons; } d0_bucket;

static int d0_cmp(const void *pa, const void *pb) {
  const d0_bucket *a = (const d0_bucket *)pa, *b = (const d0_bucket *)pb;
  if (a->mass != b->mass) return a->mass > b->mass ? -1 : 1;
  return a->first < b->first ? -1 : (a->first > b->first ? 1 : 0);
}

uint32_t orc_d0_dereplicate(const orc_db *db, uint32_t *rep, uint32_t *next, uint32_t *seeds, uint64_t *mass,
                            uint32_t *size, uint32_t *singletons) {
  const uint32_t n = db->n;
  const uint64_t slots = orc_hashtable_size(n);
  d0_bucket *tab = (d0_bucket *)calloc(slots, sizeof *tab);
  uint32_t clusters = 0;
  orc_zobrist_init(db->longest + 2);
  memset(next, 0, (size_t)n * sizeof *next);                  /* 0 terminates a chain (:399) */
  for (uint32_t a = 0; a < n; ++a) {
    const uint64_t *seq = db->words + db->off[a];
    const uint32_t len = db->len[a];
    const uint64_t h = orc_zobrist_hash(seq, len);
    uint64_t j = h & (slots - 1);
    while (tab[j].mass != 0 &&
           (tab[j].hash != h || db->len[tab[j].first] != len ||
            memcmp(seq, db->words + db->off[tab[j].first], (size_t)((len + 31) / 32) * 8) != 0))   /* nt_bytelength, :306 */
      j = (j + 1) & (slots - 1);
    if (tab[j].mass != 0) next[tab[j].last] = a;
    else { ++clusters; tab[j].hash = h; tab[j].first = a; tab[j].size = 0; tab[j].singletons = 0; }
    tab[j].size++;
    tab[j].last = a;
    tab[j].mass += db->abundance[a];
    if (db->abundance[a] == 1) tab[j].singletons++;
    rep[a] = tab[j].first;
  }
  qsort(tab, slots, sizeof *tab, d0_cmp);                      /* empty buckets (mass 0) sink to the end */
  for (uint32_t k = 0; k < clusters; ++k) {
    seeds[k] = tab[k].first; mass[k] = tab[k].mass; size[k] = tab[k].size; singletons[k] = tab[k].singletons;
  }
  free(tab);
  return clusters;
}
