/* oracle/oracle_dn.c — plain-C restatement of swarm's d>1 path (TEST INFRASTRUCTURE, see oracle.h).
 *
 * q-gram parity vectors and the popcount lower bound (src/qgram.cc:68-96,247-252), the scalar global
 * aligner with the reference's tie-breaks (src/nw.cc:40-191 — the SIMD kernels src/search8.cc /
 * src/search16.cc + src/utils/backtrack.h compute the same recurrence with inverted flag polarity),
 * the scoring conversion (src/swarm.cc:466-483) and the greedy control loop with its in-place list
 * rotations and triangle-inequality pruning (src/algo.cc:384-602, :205-256).
 * Pinned by the d2 / d3 fixtures of tests/golden/ (outputs of the unmodified reference binary, which runs the
 * SIMD path).  Citations are file:line under /root/reference.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline uint32_t nt_extract(const uint64_t *seq, uint32_t pos) {
  return (uint32_t)((seq[pos >> 5] >> ((pos & 31u) << 1)) & 3u);
}
static inline const uint64_t *db_seq(const orc_db *db, uint32_t i) { return db->words + db->off[i]; }

/* src/qgram.cc:68-96: 1024-bit parity vector of 5-mer occurrences */
void orc_findqgrams(const uint64_t *seq, uint32_t len, uint8_t *vec) {
  memset(vec, 0, 128);
  uint64_t qgram = 0;
  uint32_t pos = 0;
  while (pos < 4 && pos < len) { qgram = (qgram << 2) | nt_extract(seq, pos); pos++; }
  while (pos < len) {
    qgram = (qgram << 2) | nt_extract(seq, pos);
    vec[(qgram >> 3) & 127] ^= (uint8_t)(1u << (qgram & 7));
    pos++;
  }
}
/* src/qgram.cc:247-252 with src/popcnt.cc:45-62: ceil(popcount(a^b) / 10) */
uint64_t orc_qgram_diff(const uint8_t *a, const uint8_t *b) {
  uint64_t c = 0;
  for (int i = 0; i < 128; i++) c += (uint64_t)__builtin_popcount((unsigned)(a[i] ^ b[i]));
  return (c + 9) / 10;
}

/* src/swarm.cc:466-483: (match reward, mismatch penalty, gap open, gap extend) -> converted costs / gcd */
static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }
void orc_scoring(int64_t m, int64_t p, int64_t g, int64_t e, int64_t out[3]) {
  int64_t mis = 2 * m + 2 * p, go = 2 * g, ge = m + 2 * e;
  const int64_t f = gcd64(gcd64(mis, go), ge);
  out[0] = mis / f; out[1] = go / f; out[2] = ge / f;
}

/* src/nw.cc:40-112 (align) + :115-191 (backtrack): differences of THE optimal alignment.
 * rows = database/target sequence d, columns = query sequence q. */
uint64_t orc_nw_diffs(const uint64_t *dseq, uint32_t dlen, const uint64_t *qseq, uint32_t qlen,
                      int64_t mismatch, int64_t gapopen_, int64_t gapextend_, uint64_t *alnlen) {
  const uint64_t gapopen = (uint64_t)gapopen_, gapextend = (uint64_t)gapextend_;
  uint8_t *dir = (uint8_t *)calloc((size_t)qlen * dlen + 1, 1);
  uint64_t *he = (uint64_t *)malloc((size_t)2 * qlen * sizeof(uint64_t) + 16);
  for (uint64_t c = 0; c < qlen; c++) {
    he[2 * c] = gapopen + (c + 1) * gapextend;
    he[2 * c + 1] = 2 * gapopen + (c + 2) * gapextend;
  }
  for (uint64_t r = 0; r < dlen; r++) {
    uint64_t top = 2 * gapopen + (r + 2) * gapextend;
    uint64_t diagonal = r == 0 ? 0 : gapopen + r * gapextend;
    const uint32_t db = nt_extract(dseq, (uint32_t)r);
    for (uint64_t c = 0; c < qlen; c++) {
      const uint64_t idx = (uint64_t)qlen * r + c;
      const uint64_t prevdiag = he[2 * c];
      uint64_t left = he[2 * c + 1];
      diagonal += (db == nt_extract(qseq, (uint32_t)c)) ? 0u : (uint64_t)mismatch;
      if (top < diagonal) dir[idx] |= 1;                 /* maskup */
      if (top < diagonal) diagonal = top;
      if (left < diagonal) diagonal = left;
      if (left == diagonal) dir[idx] |= 2;               /* maskleft */
      he[2 * c] = diagonal;
      diagonal += gapopen + gapextend;
      left += gapextend;
      top += gapextend;
      if (top < diagonal) dir[idx] |= 4;                 /* maskextup */
      if (left < diagonal) dir[idx] |= 8;                /* maskextleft */
      if (diagonal < top) top = diagonal;
      if (diagonal < left) left = diagonal;
      he[2 * c + 1] = left;
      diagonal = prevdiag;
    }
  }
  uint64_t alength = 0, matches = 0, column = qlen, row = dlen;
  char op = 0;
  while (column > 0 && row > 0) {
    const uint8_t cell = dir[(uint64_t)qlen * (row - 1) + (column - 1)];
    alength++;
    if (op == 'I' && (cell & 8)) { row--; }
    else if (op == 'D' && (cell & 4)) { column--; }
    else if (cell & 2) { row--; op = 'I'; }
    else if (cell & 1) { column--; op = 'D'; }
    else {
      if (nt_extract(qseq, (uint32_t)column - 1) == nt_extract(dseq, (uint32_t)row - 1)) matches++;
      column--; row--; op = 'M';
    }
  }
  alength += column + row;
  free(dir); free(he);
  if (alnlen) *alnlen = alength;
  return alength - matches;
}

typedef struct { uint32_t ampliconid, diffestimate, swarmid, generation, radius; } ampinfo;   /* src/algo.cc:67-74 */

/* src/algo.cc:329-708 (clustering part).  Outputs, all indexed by amplicon id unless stated:
 *  order[n]        final list order (amps_v[i].ampliconid)
 *  swarm_of[id]    amplicon id of the seed of id's swarm;  generation[id] (seed 0, first hits 1, ...)
 *  parent[id]      the (sub)seed that accepted id (ORC_NONE for seeds);  pdiff[id] = differences to parent
 *  radius[id]      accumulated differences from the seed (:493, :569)
 *  stats (NULL or [3]): q-gram comparisons, alignments, accepted links.   Returns number of swarms. */
uint32_t orc_dn_cluster(const orc_db *db, uint32_t d, int ncb, const int64_t pen[3],
                        uint32_t *order, uint32_t *swarm_of, uint32_t *generation, uint32_t *parent,
                        uint32_t *pdiff, uint32_t *radius, uint64_t *stats) {
  const uint32_t n = db->n;
  uint8_t *qg = (uint8_t *)malloc((size_t)n * 128);
  for (uint32_t i = 0; i < n; i++) orc_findqgrams(db_seq(db, i), db->len[i], qg + (size_t)i * 128);   /* db_qgrams_init src/db.cc:819-842 */
  ampinfo *amps = (ampinfo *)calloc(n, sizeof(ampinfo));
  for (uint32_t i = 0; i < n; i++) amps[i].ampliconid = i;
  uint64_t *tind = (uint64_t *)malloc((size_t)n * 8), *tamp = (uint64_t *)malloc((size_t)n * 8);
  uint64_t st_q = 0, st_a = 0, st_l = 0;
  uint64_t seeded = 0, swarmed = 0;
  uint32_t swarmid = 0;
  for (uint32_t i = 0; i < n; i++) { parent[i] = ORC_NONE; pdiff[i] = 0; radius[i] = 0; generation[i] = 0; }
  while (seeded < n) {
    swarmid++;
    const uint64_t seedindex = seeded++;
    amps[seedindex].swarmid = swarmid;
    const uint32_t seedamp = amps[seedindex].ampliconid;
    swarm_of[seedamp] = seedamp;
    const uint64_t seedab = db->abundance[seedamp];
    swarmed++;
    /* diff estimates between the seed and every remaining amplicon (:416-449) */
    uint64_t targetcount = 0, listlen = 0;
    for (uint64_t i = swarmed; i < n; i++) {
      const uint32_t a = amps[i].ampliconid;
      if (ncb || db->abundance[a] <= seedab) {
        const uint64_t diff = orc_qgram_diff(qg + (size_t)seedamp * 128, qg + (size_t)a * 128);
        st_q++;
        amps[swarmed + listlen].diffestimate = (uint32_t)diff;     /* sic: indexed by list position (:441) */
        if (diff <= d) { tind[targetcount] = swarmed + listlen; tamp[targetcount] = a; targetcount++; }
        listlen++;
      }
    }
    if (targetcount == 0) continue;
    /* the reference aligns all targets first (search_do :453), then accepts in list order (:456-502) */
    uint64_t *dv = (uint64_t *)malloc(targetcount * 8);
    for (uint64_t t = 0; t < targetcount; t++) {
      dv[t] = orc_nw_diffs(db_seq(db, (uint32_t)tamp[t]), db->len[tamp[t]], db_seq(db, seedamp), db->len[seedamp], pen[0], pen[1], pen[2], NULL);
      st_a++;
    }
    for (uint64_t t = 0; t < targetcount; t++) {
      if (dv[t] > d) continue;
      const uint64_t target = tind[t];
      if (target > swarmed) {                              /* move_target_to_first_unswarmed_position :222-256 */
        const ampinfo tmp = amps[target];
        for (uint64_t i = target; i > swarmed; i--) amps[i] = amps[i - 1];
        amps[swarmed] = tmp;
      }
      amps[swarmed].swarmid = swarmid; amps[swarmed].generation = 1; amps[swarmed].radius = (uint32_t)dv[t];
      const uint32_t a = amps[swarmed].ampliconid;
      swarm_of[a] = seedamp; generation[a] = 1; parent[a] = seedamp; pdiff[a] = (uint32_t)dv[t]; radius[a] = (uint32_t)dv[t];
      st_l++;
      swarmed++;
    }
    free(dv);
    while (seeded < swarmed) {                             /* subseeds :505-602 */
      const ampinfo subseed = amps[seeded];
      seeded++;
      targetcount = 0;
      const uint64_t subab = db->abundance[subseed.ampliconid];
      for (uint64_t i = swarmed; i < n; i++) {
        const uint32_t a = amps[i].ampliconid;
        if (amps[i].diffestimate <= subseed.radius + d && (ncb || db->abundance[a] <= subab)) {
          st_q++;
          if (orc_qgram_diff(qg + (size_t)subseed.ampliconid * 128, qg + (size_t)a * 128) <= d) { tind[targetcount] = i; tamp[targetcount] = a; targetcount++; }
        }
      }
      if (targetcount == 0) continue;
      uint64_t *dv2 = (uint64_t *)malloc(targetcount * 8);
      for (uint64_t t = 0; t < targetcount; t++) {
        dv2[t] = orc_nw_diffs(db_seq(db, (uint32_t)tamp[t]), db->len[tamp[t]], db_seq(db, subseed.ampliconid), db->len[subseed.ampliconid], pen[0], pen[1], pen[2], NULL);
        st_a++;
      }
      for (uint64_t t = 0; t < targetcount; t++) {
        if (dv2[t] > d) continue;
        const uint64_t target = tind[t];
        uint64_t pos = swarmed;                            /* find_correct_position_in_list :205-219 */
        const uint32_t tid = amps[target].ampliconid;
        while (pos > seeded && amps[pos - 1].ampliconid > tid && amps[pos - 1].generation > subseed.generation) pos--;
        if (target > pos) {
          const ampinfo tmp = amps[target];
          for (uint64_t i = target; i > pos; i--) amps[i] = amps[i - 1];
          amps[pos] = tmp;
        }
        amps[pos].swarmid = swarmid; amps[pos].generation = subseed.generation + 1;
        amps[pos].radius = subseed.radius + (uint32_t)dv2[t];
        const uint32_t a = amps[pos].ampliconid;
        swarm_of[a] = seedamp; generation[a] = subseed.generation + 1; parent[a] = subseed.ampliconid;
        pdiff[a] = (uint32_t)dv2[t]; radius[a] = amps[pos].radius;
        st_l++;
        swarmed++;
      }
      free(dv2);
    }
  }
  for (uint32_t i = 0; i < n; i++) order[i] = amps[i].ampliconid;
  free(qg); free(amps); free(tind); free(tamp);
  if (stats) { stats[0] = st_q; stats[1] = st_a; stats[2] = st_l; }
  return swarmid;
}
