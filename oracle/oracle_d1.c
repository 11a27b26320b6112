/* oracle/oracle_d1.c — plain-C restatement of swarm's d=1 path (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Follows the reference ALGORITHM step by step (variant enumeration order, Bloom pre-test, bucket
 * walk, abundance rule, greedy generation-by-generation BFS, two-level fastidious search) — not the
 * closed forms the GPU engine uses — so that it is an independent check of those closed forms.
 * Citations are file:line under /root/reference.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ mt19937_64 (std::mt19937_64)
 * The reference draws all its random tables from one `std::mt19937_64 rand_64(1)`
 * (src/utils/pseudo_rng.h:30-31).  This is the published MT19937-64 recurrence (Matsumoto &
 * Nishimura 2004; C++11 [rand.predef]: w=64 n=312 m=156 r=31 a=0xB5026F5AA96619E9 u=29
 * d=0x5555555555555555 s=17 b=0x71D67FFFEDA60000 t=37 c=0xFFF7EEE000000000 l=43 f=6364136223846793005). */
static uint64_t mt[312];
static int mti = 313;
static void mt_seed(uint64_t seed) {
  mt[0] = seed;
  for (mti = 1; mti < 312; mti++) mt[mti] = 6364136223846793005ULL * (mt[mti - 1] ^ (mt[mti - 1] >> 62)) + (uint64_t)mti;
}
void orc_mt19937_64_seed(uint64_t seed) { mt_seed(seed); }
uint64_t orc_mt19937_64_next(void) {
  static const uint64_t mag01[2] = {0ULL, 0xB5026F5AA96619E9ULL};
  if (mti >= 312) {
    if (mti == 313) mt_seed(5489ULL);
    int i;
    for (i = 0; i < 312 - 156; i++) {
      uint64_t x = (mt[i] & 0xFFFFFFFF80000000ULL) | (mt[i + 1] & 0x7FFFFFFFULL);
      mt[i] = mt[i + 156] ^ (x >> 1) ^ mag01[x & 1ULL];
    }
    for (; i < 311; i++) {
      uint64_t x = (mt[i] & 0xFFFFFFFF80000000ULL) | (mt[i + 1] & 0x7FFFFFFFULL);
      mt[i] = mt[i + (156 - 312)] ^ (x >> 1) ^ mag01[x & 1ULL];
    }
    uint64_t x = (mt[311] & 0xFFFFFFFF80000000ULL) | (mt[0] & 0x7FFFFFFFULL);
    mt[311] = mt[155] ^ (x >> 1) ^ mag01[x & 1ULL];
    mti = 0;
  }
  uint64_t x = mt[mti++];
  x ^= (x >> 29) & 0x5555555555555555ULL;
  x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
  x ^= (x << 37) & 0xFFF7EEE000000000ULL;
  x ^= (x >> 43);
  return x;
}

/* ------------------------------------------------------------------ 2-bit codec
 * src/utils/nt_codec.cc:35-62 (nt_extract), :65-75 (nt_bytelength); src/variants.cc:33-45 (nt_set) */
static inline uint32_t nt_extract(const uint64_t *seq, uint32_t pos) {
  return (uint32_t)((seq[pos >> 5] >> ((pos & 31u) << 1)) & 3u);
}
static inline void nt_set(uint64_t *seq, uint32_t pos, uint32_t base) {
  const uint32_t sh = (pos & 31u) << 1;
  seq[pos >> 5] = (seq[pos >> 5] & ~(3ULL << sh)) | ((uint64_t)base << sh);
}
static inline uint32_t nt_words(uint32_t len) { return (len + 31u) >> 5; }

/* ------------------------------------------------------------------ Zobrist
 * src/zobrist.cc:49-80: value = four rounds of (x << 16) ^ rand_64(); table index 4*pos + base */
static uint64_t *ztab = NULL;
static uint32_t zlen = 0;
void orc_zobrist_init(uint32_t zobrist_len) {
  free(ztab);
  zlen = zobrist_len;
  ztab = (uint64_t *)malloc(4ULL * zobrist_len * sizeof(uint64_t));
  mt_seed(1);
  for (uint64_t i = 0; i < 4ULL * zobrist_len; i++) {
    uint64_t v = orc_mt19937_64_next();
    v <<= 16; v ^= orc_mt19937_64_next();
    v <<= 16; v ^= orc_mt19937_64_next();
    v <<= 16; v ^= orc_mt19937_64_next();
    ztab[i] = v;
  }
}
void orc_zobrist_exit(void) { free(ztab); ztab = NULL; zlen = 0; }
uint64_t orc_zobrist_value(uint32_t pos, uint32_t base) { return ztab[4ULL * pos + base]; }  /* :127-131 */
/* src/zobrist.cc:134-184 computes XOR_p Z[p][s_p] through a byte-combined table (:83-108); the
 * combination is exact, so the plain per-position XOR below returns the same value. */
uint64_t orc_zobrist_hash(const uint64_t *seq, uint32_t len) {
  uint64_t h = 0;
  for (uint32_t p = 0; p < len; p++) h ^= orc_zobrist_value(p, nt_extract(seq, p));
  return h;
}
static uint64_t zobrist_hash_delete_first(const uint64_t *seq, uint32_t len) {   /* :189-213 */
  uint64_t h = 0;
  for (uint32_t p = 1; p < len; p++) h ^= orc_zobrist_value(p - 1, nt_extract(seq, p));
  return h;
}
static uint64_t zobrist_hash_insert_first(const uint64_t *seq, uint32_t len) {   /* :216-240 */
  uint64_t h = 0;
  for (uint32_t p = 0; p < len; p++) h ^= orc_zobrist_value(p + 1, nt_extract(seq, p));
  return h;
}

/* ------------------------------------------------------------------ microvariants */
static inline void add_variant(orc_var *out, uint32_t *cnt, uint64_t hash, uint8_t type, uint32_t pos, uint8_t base) {
  orc_var *v = &out[(*cnt)++];
  v->hash = hash; v->type = type; v->pos = pos; v->base = base; v->pad = 0;
}
/* src/variants.cc:184-249 — same enumeration order, same canonicalisation */
uint32_t orc_generate_variants(const uint64_t *seq, uint32_t len, uint64_t hash, orc_var *out) {
  uint32_t cnt = 0;
  for (uint32_t p = 0; p < len; p++) {                                /* substitutions :192-206 */
    const uint32_t cur = nt_extract(seq, p);
    const uint64_t h1 = hash ^ orc_zobrist_value(p, cur);
    for (uint32_t b = 0; b < 4; b++) {
      if (b == cur) continue;
      add_variant(out, &cnt, h1 ^ orc_zobrist_value(p, b), 0, p, (uint8_t)b);
    }
  }
  uint64_t h = zobrist_hash_delete_first(seq, len);                   /* deletions :210-222 */
  add_variant(out, &cnt, h, 1, 0, 0);
  uint32_t prev = nt_extract(seq, 0);
  for (uint32_t p = 1; p < len; p++) {
    const uint32_t cur = nt_extract(seq, p);
    if (cur == prev) continue;
    h ^= orc_zobrist_value(p - 1, prev) ^ orc_zobrist_value(p - 1, cur);
    add_variant(out, &cnt, h, 1, p, 0);
    prev = cur;
  }
  h = zobrist_hash_insert_first(seq, len);                            /* insertions :226-246 */
  for (uint32_t b = 0; b < 4; b++) add_variant(out, &cnt, h ^ orc_zobrist_value(0, b), 2, 0, (uint8_t)b);
  for (uint32_t p = 0; p < len; p++) {
    const uint32_t cur = nt_extract(seq, p);
    h ^= orc_zobrist_value(p, cur) ^ orc_zobrist_value(p + 1, cur);
    for (uint32_t b = 0; b < 4; b++) {
      if (b == cur) continue;
      add_variant(out, &cnt, h ^ orc_zobrist_value(p + 1, b), 2, p + 1, (uint8_t)b);
    }
  }
  return cnt;
}

static int seq_identical(const uint64_t *a, uint32_t as, const uint64_t *b, uint32_t bs, uint32_t length) {
  for (uint32_t i = 0; i < length; i++)                                /* src/variants.cc:62-75 */
    if (nt_extract(a, as + i) != nt_extract(b, bs + i)) return 0;
  return 1;
}
/* src/variants.cc:118-165 */
int orc_check_variant(const uint64_t *seed, uint32_t seedlen, const orc_var *v, const uint64_t *amp, uint32_t amplen) {
  switch (v->type) {
    case 0:
      return seedlen == amplen && seq_identical(seed, 0, amp, 0, v->pos) &&
             nt_extract(amp, v->pos) == v->base &&
             seq_identical(seed, v->pos + 1, amp, v->pos + 1, seedlen - v->pos - 1);
    case 1:
      return seedlen - 1 == amplen && seq_identical(seed, 0, amp, 0, v->pos) &&
             seq_identical(seed, v->pos + 1, amp, v->pos, seedlen - v->pos - 1);
    default:
      return seedlen + 1 == amplen && seq_identical(seed, 0, amp, 0, v->pos) &&
             nt_extract(amp, v->pos) == v->base &&
             seq_identical(seed, v->pos, amp, v->pos + 1, seedlen - v->pos);
  }
}
/* src/variants.cc:78-115 */
void orc_generate_variant_sequence(const uint64_t *seed, uint32_t seedlen, const orc_var *v, uint64_t *out, uint32_t *outlen) {
  uint32_t i;
  switch (v->type) {
    case 0:
      memcpy(out, seed, nt_words(seedlen) * 8u);
      nt_set(out, v->pos, v->base);
      *outlen = seedlen;
      break;
    case 1:
      for (i = 0; i < v->pos; i++) nt_set(out, i, nt_extract(seed, i));
      for (i = 0; i < seedlen - v->pos - 1; i++) nt_set(out, v->pos + i, nt_extract(seed, v->pos + 1 + i));
      *outlen = seedlen - 1;
      break;
    default:
      for (i = 0; i < v->pos; i++) nt_set(out, i, nt_extract(seed, i));
      nt_set(out, v->pos, v->base);
      for (i = 0; i < seedlen - v->pos; i++) nt_set(out, v->pos + 1 + i, nt_extract(seed, v->pos + i));
      *outlen = seedlen + 1;
      break;
  }
}

/* ------------------------------------------------------------------ hash table + Bloom filters */
/* src/utils/hashtable_size.cc:29-42: smallest power of two >= 10(n+1)/7 (integer division first) */
uint64_t orc_hashtable_size(uint64_t n) {
  const double x = (double)(10 * (n + 1) / 7);
  return (uint64_t)pow(2.0, ceil(log(x) / log(2.0)));
}

typedef struct {                     /* src/hashtable.cc:41-44,125-146: SoA, linear probing */
  uint64_t mask; uint8_t *occupied; uint64_t *values; uint32_t *data;
} htab;
static inline uint64_t h_index(const htab *t, uint64_t hash) { return (hash >> 32) & t->mask; }        /* :47-53 */
static inline int h_occ(const htab *t, uint64_t i) { return (t->occupied[i >> 3] >> (i & 7)) & 1; }    /* :76-87 */
static inline void h_setocc(htab *t, uint64_t i) { t->occupied[i >> 3] |= (uint8_t)(1u << (i & 7)); }  /* :62-73 */

typedef struct { uint64_t mask; uint64_t *bitmap; uint64_t patterns[1024]; } bloompat;   /* src/bloompat.h:28-39 */
static void bloom_patterns(uint64_t *pat, uint32_t count, uint32_t k) {      /* src/bloompat.cc:74-90 */
  for (uint32_t i = 0; i < count; i++) {
    uint64_t p = 0;
    for (uint32_t j = 0; j < k; j++) {
      uint64_t one = 1ULL << (orc_mt19937_64_next() & 63u);
      while (p & one) one = 1ULL << (orc_mt19937_64_next() & 63u);
      p |= one;
    }
    pat[i] = p;
  }
}
static void bloom_init(bloompat *b, uint64_t size_bytes) {                   /* src/bloompat.cc:100-120 */
  if (size_bytes < 8) size_bytes = 8;
  b->mask = (size_bytes >> 3) - 1;
  b->bitmap = (uint64_t *)malloc(size_bytes);
  memset(b->bitmap, 0xFF, size_bytes);
  bloom_patterns(b->patterns, 1024, 8);
}
static inline void bloom_set(bloompat *b, uint64_t h) { b->bitmap[(h >> 10) & b->mask] &= ~b->patterns[h & 1023]; }   /* :62-65 */
static inline int bloom_get(const bloompat *b, uint64_t h) { return (b->bitmap[(h >> 10) & b->mask] & b->patterns[h & 1023]) == 0; } /* :68-71 */

typedef struct { uint64_t size; uint64_t *bitmap; uint64_t *patterns; } bloomflex;       /* src/bloomflex.h:28-39 */
static void bloomflex_init(bloomflex *b, uint64_t bytes, uint32_t k) {       /* src/bloomflex.cc:93-115 */
  b->size = bytes >> 3;
  b->patterns = (uint64_t *)malloc(65536 * sizeof(uint64_t));
  bloom_patterns(b->patterns, 65536, k);
  b->bitmap = (uint64_t *)malloc(b->size * 8);
  memset(b->bitmap, 0xFF, b->size * 8);
}
static inline void bloomflex_set(bloomflex *b, uint64_t h) { b->bitmap[(h >> 16) % b->size] &= ~b->patterns[h & 65535]; } /* :61-64 */
static inline int bloomflex_get(const bloomflex *b, uint64_t h) { return (b->bitmap[(h >> 16) % b->size] & b->patterns[h & 65535]) == 0; } /* :67-70 */

static inline const uint64_t *db_seq(const orc_db *db, uint32_t i) { return db->words + db->off[i]; }

/* src/algod1.cc:174-208: insert, flag identical sequences */
static int hash_insert(const orc_db *db, htab *t, bloompat *bl, const uint64_t *hashes, uint32_t amp) {
  const uint64_t hash = hashes[amp];
  uint64_t idx = h_index(t, hash);
  int dup = 0;
  while (h_occ(t, idx)) {
    if (t->values[idx] == hash) {
      const uint32_t other = t->data[idx];
      if (db->len[other] == db->len[amp] &&
          memcmp(db_seq(db, other), db_seq(db, amp), nt_words(db->len[amp]) * 8u) == 0) dup = 1;
    }
    idx = (idx + 1) & t->mask;
  }
  h_setocc(t, idx);
  t->values[idx] = hash;
  t->data[idx] = amp;
  bloom_set(bl, hash);
  return dup;
}

static void htab_alloc(htab *t, uint64_t n) {
  const uint64_t sz = orc_hashtable_size(n);
  t->mask = sz - 1;
  t->occupied = (uint8_t *)calloc((sz + 63) / 8, 1);
  t->values = (uint64_t *)malloc(sz * sizeof(uint64_t));
  t->data = (uint32_t *)malloc(sz * sizeof(uint32_t));
}
static void htab_free(htab *t) { free(t->occupied); free(t->values); free(t->data); }

/* ------------------------------------------------------------------ d=1 network */
int orc_d1_network(const orc_db *db, int no_cluster_breaking, uint32_t *link_start, uint32_t *link_count,
                   uint32_t **network, uint64_t *n_edges, uint64_t *stats) {
  const uint32_t n = db->n;
  uint64_t st_var = 0, st_bloom = 0, st_slots = 0, st_cmp = 0;
  orc_zobrist_init(db->longest + 2);                       /* src/db.cc:652-653 (header term irrelevant) */
  uint64_t *hashes = (uint64_t *)malloc((size_t)n * 8);
  for (uint32_t i = 0; i < n; i++) hashes[i] = orc_zobrist_hash(db_seq(db, i), db->len[i]);   /* src/db.cc:761 */

  htab t; htab_alloc(&t, n);                               /* src/algod1.cc:1119-1127 */
  bloompat bl; bloom_init(&bl, t.mask + 1);
  int dup = 0;
  for (uint32_t k = 0; k < n && !dup; k++) dup = hash_insert(db, &t, &bl, hashes, k);         /* :1132-1139 */
  if (dup) { htab_free(&t); free(bl.bitmap); free(hashes); *network = NULL; *n_edges = 0; return 1; }

  orc_var *vars = (orc_var *)malloc((7ULL * db->longest + 5) * sizeof(orc_var));
  uint64_t cap = 1u << 20, cnt = 0;
  uint32_t *net = (uint32_t *)malloc(cap * sizeof(uint32_t));
  for (uint32_t seed = 0; seed < n; seed++) {              /* network_thread :630-670, check_variants :606-627 */
    const uint64_t *sseq = db_seq(db, seed);
    const uint32_t slen = db->len[seed];
    const uint32_t nv = orc_generate_variants(sseq, slen, hashes[seed], vars);
    st_var += nv;
    link_start[seed] = (uint32_t)cnt;
    uint32_t hits = 0;
    for (uint32_t i = 0; i < nv; i++) {                    /* find_variant_matches :558-603 */
      const orc_var *v = &vars[i];
      if (!bloom_get(&bl, v->hash)) continue;
      st_bloom++;
      uint64_t idx = h_index(&t, v->hash);
      while (h_occ(&t, idx)) {
        st_slots++;
        if (t.values[idx] == v->hash) {
          const uint32_t amp = t.data[idx];
          if (seed != amp && (no_cluster_breaking || db->abundance[seed] >= db->abundance[amp])) {   /* :580-582 */
            st_cmp++;
            if (orc_check_variant(sseq, slen, v, db_seq(db, amp), db->len[amp])) {
              if (cnt + 1 > cap) { cap *= 2; net = (uint32_t *)realloc(net, cap * sizeof(uint32_t)); }
              net[cnt++] = amp;
              hits++;
              break;                                       /* :596 */
            }
          }
        }
        idx = (idx + 1) & t.mask;
      }
    }
    link_count[seed] = hits;
  }
  free(vars); htab_free(&t); free(bl.bitmap); free(hashes);
  *network = net; *n_edges = cnt;
  if (stats) { stats[0] = st_var; stats[1] = st_bloom; stats[2] = st_slots; stats[3] = st_cmp; }
  return 0;
}

/* ------------------------------------------------------------------ d=1 greedy clustering */
static int cmp_u32(const void *a, const void *b) {
  const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : (x > y);
}

typedef struct {
  const orc_db *db; const uint32_t *ls, *lc, *net;
  uint32_t *swarmid, *generation, *parent;
  uint32_t *hits; uint64_t hits_cap; uint32_t hits_cnt;
  uint32_t swarmsize, swarm_maxgen, singletons; uint64_t mass, sumlen;
} bfs_state;

/* process_seed src/algod1.cc:673-718 */
static void process_seed(bfs_state *s, uint32_t seed) {
  s->swarmsize++;
  if (s->generation[seed] > s->swarm_maxgen) s->swarm_maxgen = s->generation[seed];
  const uint64_t ab = s->db->abundance[seed];
  s->mass += ab;
  if (ab == 1) s->singletons++;
  s->sumlen += s->db->len[seed];
  const uint32_t start = s->ls[seed], count = s->lc[seed];
  if ((uint64_t)s->hits_cnt + count > s->hits_cap) {
    while ((uint64_t)s->hits_cnt + count > s->hits_cap) s->hits_cap += 4096;
    s->hits = (uint32_t *)realloc(s->hits, s->hits_cap * sizeof(uint32_t));
  }
  for (uint32_t o = 0; o < count; o++) {
    const uint32_t amp = s->net[start + o];
    if (s->swarmid[amp] == ORC_NONE) {
      s->hits[s->hits_cnt++] = amp;
      s->swarmid[amp] = s->swarmid[seed];
      s->generation[amp] = s->generation[seed] + 1;
      s->parent[amp] = seed;
    }
  }
}

uint32_t orc_d1_cluster(const orc_db *db, const uint32_t *link_start, const uint32_t *link_count, const uint32_t *network,
                        uint32_t *swarmid, uint32_t *generation, uint32_t *parent, uint32_t *next,
                        uint32_t *sw_seed, uint32_t *sw_last, uint32_t *sw_size, uint32_t *sw_singletons,
                        uint32_t *sw_maxgen, uint64_t *sw_mass, uint64_t *sw_sumlen) {
  const uint32_t n = db->n;
  bfs_state s;
  memset(&s, 0, sizeof s);
  s.db = db; s.ls = link_start; s.lc = link_count; s.net = network;
  s.swarmid = swarmid; s.generation = generation; s.parent = parent;
  s.hits_cap = 7ULL * db->longest + 5;
  s.hits = (uint32_t *)malloc(s.hits_cap * sizeof(uint32_t));
  for (uint32_t i = 0; i < n; i++) { swarmid[i] = ORC_NONE; parent[i] = 0; generation[i] = 0; next[i] = ORC_NONE; }  /* ampinfo_s defaults :87-96 */
  uint32_t swarmcount = 0;
  for (uint32_t seed = 0; seed < n; seed++) {              /* src/algod1.cc:1185-1280 */
    if (swarmid[seed] != ORC_NONE) continue;
    swarmid[seed] = swarmcount; generation[seed] = 0; parent[seed] = ORC_NONE; next[seed] = ORC_NONE;
    uint32_t tail = seed;
    s.swarmsize = 0; s.swarm_maxgen = 0; s.mass = 0; s.singletons = 0; s.sumlen = 0;
    s.hits_cnt = 0;
    process_seed(&s, seed);
    qsort(s.hits, s.hits_cnt, sizeof(uint32_t), cmp_u32);
    for (uint32_t i = 0; i < s.hits_cnt; i++) { next[tail] = s.hits[i]; tail = s.hits[i]; }   /* add_amp_to_swarm :746-752 */
    uint32_t subseed = next[seed];
    while (subseed != ORC_NONE) {
      s.hits_cnt = 0;
      while (subseed != ORC_NONE) { process_seed(&s, subseed); subseed = next[subseed]; }
      qsort(s.hits, s.hits_cnt, sizeof(uint32_t), cmp_u32);
      for (uint32_t i = 0; i < s.hits_cnt; i++) { next[tail] = s.hits[i]; tail = s.hits[i]; }
      subseed = s.hits_cnt ? s.hits[0] : ORC_NONE;
    }
    sw_seed[swarmcount] = seed; sw_size[swarmcount] = s.swarmsize; sw_mass[swarmcount] = s.mass;
    sw_sumlen[swarmcount] = s.sumlen; sw_singletons[swarmcount] = s.singletons;
    sw_maxgen[swarmcount] = s.swarm_maxgen; sw_last[swarmcount] = tail;
    swarmcount++;
  }
  free(s.hits);
  return swarmcount;
}

/* ------------------------------------------------------------------ fastidious */
typedef struct { uint32_t parent, child; } graft_pair;
static int cmp_graft(const void *a, const void *b) {       /* src/algod1.cc:297-309 */
  const graft_pair *x = (const graft_pair *)a, *y = (const graft_pair *)b;
  if (x->parent != y->parent) return x->parent < y->parent ? -1 : 1;
  return x->child < y->child ? -1 : (x->child > y->child);
}

int64_t orc_d1_fastidious(const orc_db *db, uint64_t boundary, uint32_t bloom_bits, uint32_t nswarms,
                          const uint32_t *swarmid, uint32_t *next,
                          uint32_t *sw_seed, uint32_t *sw_last, uint32_t *sw_size, uint32_t *sw_singletons,
                          uint64_t *sw_mass, uint64_t *sw_sumlen, uint8_t *sw_attached,
                          uint32_t *graft_cand, uint32_t *graft_raw, uint64_t *stats) {
  const uint32_t n = db->n;
  (void)sw_seed;
  for (uint32_t i = 0; i < n; i++) graft_cand[i] = ORC_NONE;
  for (uint32_t i = 0; i < nswarms; i++) sw_attached[i] = 0;
  uint64_t small = 0, amps_small = 0, nt_small = 0;        /* :1307-1321 */
  for (uint32_t i = 0; i < nswarms; i++)
    if (sw_mass[i] < boundary) { amps_small += sw_size[i]; nt_small += sw_sumlen[i]; small++; }
  const uint64_t amps_large = n - amps_small, large = nswarms - small;
  if (small == 0 || large == 0) return -1;                 /* :1330-1334 */

  uint32_t k = (uint32_t)(0.4 * bloom_bits);               /* :1355 */
  if (k < 1) k = 1;
  uint64_t bits = nt_small * 7 * bloom_bits;               /* :1357 */
  if (bits < 64) bits = 64;
  const uint64_t nbytes = (bits - 1) / 8 + 1;              /* :1401 */

  orc_zobrist_init(db->longest + 2);
  uint64_t *hashes = (uint64_t *)malloc((size_t)n * 8);
  for (uint32_t i = 0; i < n; i++) hashes[i] = orc_zobrist_hash(db_seq(db, i), db->len[i]);
  htab t; htab_alloc(&t, n);                               /* emptied table :1411 */
  bloompat bl; bloom_init(&bl, t.mask + 1);                /* zapped bloom_a :1412 */
  bloomflex bf; bloomflex_init(&bf, nbytes, k);

  orc_var *v1 = (orc_var *)malloc((7ULL * db->longest + 5) * sizeof(orc_var));
  orc_var *v2 = (orc_var *)malloc((7ULL * (db->longest + 1) + 5) * sizeof(orc_var));
  uint64_t *varseq = (uint64_t *)calloc(nt_words(db->longest + 2) + 1, 8);
  uint64_t light_variants = 0, heavy_variants = 0, candidates = 0, bf_pass = 0;

  /* light pass, least abundant first: mark_light_thread :521-552, mark_light_var :495-518 */
  uint64_t done = 0;
  for (uint32_t a = n; a-- > 0 && done < amps_small;) {
    if (sw_mass[swarmid[a]] >= boundary) continue;
    done++;
    hash_insert(db, &t, &bl, hashes, a);
    const uint32_t nv = orc_generate_variants(db_seq(db, a), db->len[a], hashes[a], v1);
    for (uint32_t i = 0; i < nv; i++) bloomflex_set(&bf, v1[i].hash);
    light_variants += nv;
  }
  /* heavy pass, most abundant first: check_heavy_thread :453-492, check_heavy_var :398-450 */
  done = 0;
  for (uint32_t h = 0; h < n && done < amps_large; h++) {
    if (sw_mass[swarmid[h]] < boundary) continue;
    done++;
    const uint64_t *hseq = db_seq(db, h);
    const uint32_t hlen = db->len[h];
    const uint32_t nv = orc_generate_variants(hseq, hlen, hashes[h], v1);
    heavy_variants += nv;
    for (uint32_t i = 0; i < nv; i++) {
      if (!bloomflex_get(&bf, v1[i].hash)) continue;
      bf_pass++;
      uint32_t vlen = 0;
      orc_generate_variant_sequence(hseq, hlen, &v1[i], varseq, &vlen);
      const uint64_t vh = orc_zobrist_hash(varseq, vlen);  /* check_heavy_var_2 :374-395 */
      const uint32_t nv2 = orc_generate_variants(varseq, vlen, vh, v2);
      for (uint32_t j = 0; j < nv2; j++) {
        if (!bloom_get(&bl, v2[j].hash)) continue;
        uint64_t idx = h_index(&t, v2[j].hash);            /* hash_check_attach :339-371 */
        while (h_occ(&t, idx)) {
          if (t.values[idx] == v2[j].hash) {
            const uint32_t amp = t.data[idx];
            if (orc_check_variant(varseq, vlen, &v2[j], db_seq(db, amp), db->len[amp])) {
              candidates++;                                /* add_graft_candidate :244-258 */
              if (graft_cand[amp] == ORC_NONE || graft_cand[amp] > h) graft_cand[amp] = h;
              break;
            }
          }
          idx = (idx + 1) & t.mask;
        }
      }
    }
  }
  if (graft_raw) memcpy(graft_raw, graft_cand, (size_t)n * sizeof(uint32_t));
  /* attach_candidates :274-336, attach :214-241 */
  uint32_t pairs = 0;
  for (uint32_t i = 0; i < n; i++) if (graft_cand[i] != ORC_NONE) pairs++;
  graft_pair *ga = (graft_pair *)malloc((pairs ? pairs : 1) * sizeof(graft_pair));
  uint32_t tk = 0;
  for (uint32_t i = 0; i < n; i++) if (graft_cand[i] != ORC_NONE) { ga[tk].parent = graft_cand[i]; ga[tk].child = i; tk++; }
  qsort(ga, pairs, sizeof(graft_pair), cmp_graft);
  int64_t grafts = 0;
  for (uint32_t i = 0; i < pairs; i++) {
    const uint32_t par = ga[i].parent, child = ga[i].child;
    const uint32_t ls = swarmid[child], hs = swarmid[par];
    if (sw_attached[ls]) { graft_cand[child] = ORC_NONE; continue; }
    next[sw_last[hs]] = sw_seed[ls];
    sw_last[hs] = sw_last[ls];
    sw_size[hs] += sw_size[ls]; sw_singletons[hs] += sw_singletons[ls];
    sw_mass[hs] += sw_mass[ls]; sw_sumlen[hs] += sw_sumlen[ls];
    sw_attached[ls] = 1;
    grafts++;
  }
  free(ga); free(v1); free(v2); free(varseq); free(hashes);
  htab_free(&t); free(bl.bitmap); free(bf.bitmap); free(bf.patterns);
  if (stats) { stats[0] = light_variants; stats[1] = heavy_variants; stats[2] = candidates; stats[3] = bf_pass; }
  return grafts;
}

void orc_free(void *p) { free(p); }
