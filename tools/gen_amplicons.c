/* tools/gen_amplicons.c — deterministic synthetic amplicon sets (bench/test infrastructure).
 *
 * Implements the generator specified in BASELINE.md §3.2 / SURVEY.md §8(d):
 *   until N unique sequences exist:
 *     centroid  = L uniform-random nt, abundance U[50, 5000]
 *     cluster   = max(1, floor(Exp(mean 20))) sequences, grown by picking a random existing member and
 *                 applying ONE random edit (80 % substitution / 10 % deletion / 10 % insertion),
 *                 child abundance = max(1, floor(parent * U(0, 0.5)))
 *     with probability orphan_p (0.2) emit instead an orphan = TWO edits from a member,
 *                 abundance in {1, 1, 2}   (exercises --fastidious)
 *     duplicate sequences are rejected globally (64-bit hash set; a hash collision only rejects a
 *     non-duplicate, it can never admit a duplicate)
 *   records are kept in generation order; headers are "s<i>_<abundance>".
 *
 * ab_mode 1 = "tie-heavy" variant: all abundances drawn from {1,1,1,1,1,1,1,2,3,10}
 * so that most edges are equal-abundance (bidirectional).
 *
 * Nothing here is taken from the reference; it has no generator.  Pure C11, no dependencies.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t n;          /* records generated */
  uint64_t total_nt;   /* sum of lengths */
  uint64_t *off;       /* n+1 offsets into nt */
  uint64_t *ab;        /* n abundances */
  char *nt;            /* ACGT blob */
  uint64_t cap_nt;
} gen_set;

static uint64_t s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(void) {               /* xoshiro256** */
  const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return r;
}
static void rng_seed(uint64_t x) {             /* splitmix64 */
  for (int i = 0; i < 4; i++) {
    x += 0x9e3779b97f4a7c15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    s[i] = z ^ (z >> 31);
  }
}
static inline uint64_t rng_below(uint64_t n) { return (uint64_t)(((__uint128_t)rng_next() * n) >> 64); }
static inline double rng_unit(void) { return (double)(rng_next() >> 11) * (1.0 / 9007199254740992.0); }

static uint64_t hash_bytes(const char *p, uint32_t len) {
  uint64_t h = 0x9e3779b97f4a7c15ULL ^ len;
  for (uint32_t i = 0; i < len; i++) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ULL; h ^= h >> 29; }
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ULL; h ^= h >> 32;
  return h ? h : 1;
}

static const char NT[4] = {'A', 'C', 'G', 'T'};

/* apply one random edit to src (len) -> dst; returns new length */
static uint32_t one_edit(const char *src, uint32_t len, char *dst) {
  const double u = rng_unit();
  if (u < 0.8 || len < 2) {                         /* substitution (also when too short to delete) */
    if (u >= 0.9 && len < 2) goto insertion;
    uint32_t p = (uint32_t)rng_below(len);
    memcpy(dst, src, len);
    char c;
    do { c = NT[rng_below(4)]; } while (c == src[p]);
    dst[p] = c;
    return len;
  }
  if (u < 0.9) {                                    /* deletion */
    uint32_t p = (uint32_t)rng_below(len);
    memcpy(dst, src, p);
    memcpy(dst + p, src + p + 1, len - p - 1);
    return len - 1;
  }
insertion: {
    uint32_t p = (uint32_t)rng_below((uint64_t)len + 1);
    memcpy(dst, src, p);
    dst[p] = NT[rng_below(4)];
    memcpy(dst + p + 1, src + p, len - p);
    return len + 1;
  }
}

static const uint64_t TIE_AB[10] = {1, 1, 1, 1, 1, 1, 1, 2, 3, 10};

gen_set *gen_create(uint64_t n, uint32_t L, uint64_t seed, int ab_mode, double orphan_p) {
  if (n == 0 || L == 0) return NULL;
  rng_seed(seed);
  gen_set *g = (gen_set *)calloc(1, sizeof(gen_set));
  g->off = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
  g->ab = (uint64_t *)malloc(n * sizeof(uint64_t));
  g->cap_nt = n * ((uint64_t)L + 4) + 4096;
  g->nt = (char *)malloc(g->cap_nt);
  uint64_t setcap = 1;
  while (setcap < n * 5 / 2 + 16) setcap <<= 1;
  uint64_t *set = (uint64_t *)calloc(setcap, sizeof(uint64_t));
  const uint32_t maxlen = 2 * L + 64;
  char *tmp = (char *)malloc(maxlen + 8), *tmp2 = (char *)malloc(maxlen + 8);
  uint64_t cnt = 0, pos = 0;
  g->off[0] = 0;

#define TRY_EMIT(SEQ, LEN, AB, OK)                                                        \
  do {                                                                                    \
    uint64_t h_ = hash_bytes((SEQ), (LEN)), i_ = h_ & (setcap - 1);                       \
    (OK) = 1;                                                                             \
    while (set[i_]) { if (set[i_] == h_) { (OK) = 0; break; } i_ = (i_ + 1) & (setcap - 1); } \
    if (OK) {                                                                             \
      set[i_] = h_;                                                                       \
      if (pos + (LEN) > g->cap_nt) { g->cap_nt = g->cap_nt * 5 / 4 + (LEN); g->nt = (char *)realloc(g->nt, g->cap_nt); } \
      memcpy(g->nt + pos, (SEQ), (LEN));                                                  \
      pos += (LEN); g->ab[cnt] = (AB); cnt++; g->off[cnt] = pos;                          \
    }                                                                                     \
  } while (0)

  while (cnt < n) {
    /* centroid */
    for (uint32_t i = 0; i < L; i++) tmp[i] = NT[rng_below(4)];
    uint64_t cab = ab_mode == 1 ? TIE_AB[rng_below(10)] : 50 + rng_below(4951);
    int ok;
    TRY_EMIT(tmp, L, cab, ok);
    if (!ok) continue;
    const uint64_t first = cnt - 1;          /* members of this cluster: indices first .. (members) */
    uint64_t members = 1;                    /* non-orphan members are contiguous? no: keep list */
    double e = -20.0 * log(1.0 - rng_unit());
    uint64_t want = e < 1.0 ? 1 : (uint64_t)e;
    /* member index list (generation indices); orphans are not members */
    uint64_t *mem = (uint64_t *)malloc((want + 1) * sizeof(uint64_t));
    mem[0] = first;
    uint64_t produced = 1, attempts = 0;
    while (produced < want && cnt < n && attempts < 64 * want) {
      attempts++;
      const uint64_t par = mem[rng_below(members)];
      const char *ps = g->nt + g->off[par];
      const uint32_t pl = (uint32_t)(g->off[par + 1] - g->off[par]);
      if (pl + 2 > maxlen) continue;
      if (rng_unit() < orphan_p) {
        uint32_t l1 = one_edit(ps, pl, tmp2);
        uint32_t l2 = one_edit(tmp2, l1, tmp);
        static const uint64_t OAB[3] = {1, 1, 2};
        uint64_t oab = OAB[rng_below(3)];
        TRY_EMIT(tmp, l2, oab, ok);
        if (ok) produced++;
      } else {
        uint32_t l1 = one_edit(ps, pl, tmp);
        uint64_t cab2;
        if (ab_mode == 1) cab2 = TIE_AB[rng_below(10)];
        else { cab2 = (uint64_t)((double)g->ab[par] * (rng_unit() * 0.5)); if (cab2 < 1) cab2 = 1; }
        TRY_EMIT(tmp, l1, cab2, ok);
        if (ok) { mem[members++] = cnt - 1; produced++; }
      }
    }
    free(mem);
  }
#undef TRY_EMIT
  g->n = cnt;
  g->total_nt = pos;
  free(set); free(tmp); free(tmp2);
  return g;
}

uint64_t gen_count(const gen_set *g) { return g->n; }
uint64_t gen_total_nt(const gen_set *g) { return g->total_nt; }
const char *gen_nt(const gen_set *g) { return g->nt; }
const uint64_t *gen_offsets(const gen_set *g) { return g->off; }
const uint64_t *gen_abundances(const gen_set *g) { return g->ab; }

/* copy out: nt (total_nt bytes), off (n+1), ab (n) — any pointer may be NULL */
void gen_export(const gen_set *g, char *nt, uint64_t *off, uint64_t *ab) {
  if (nt) memcpy(nt, g->nt, g->total_nt);
  if (off) memcpy(off, g->off, (g->n + 1) * sizeof(uint64_t));
  if (ab) memcpy(ab, g->ab, g->n * sizeof(uint64_t));
}

/* write ">s<i>_<ab>\n<SEQ>\n"; records [lo, hi) */
int gen_write_fasta(const gen_set *g, const char *path, uint64_t lo, uint64_t hi) {
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  static char buf[1 << 20];
  setvbuf(f, buf, _IOFBF, sizeof buf);
  if (hi > g->n) hi = g->n;
  for (uint64_t i = lo; i < hi; i++) {
    fprintf(f, ">s%llu_%llu\n", (unsigned long long)i, (unsigned long long)g->ab[i]);
    fwrite(g->nt + g->off[i], 1, g->off[i + 1] - g->off[i], f);
    fputc('\n', f);
  }
  return fclose(f);
}

void gen_free(gen_set *g) {
  if (!g) return;
  free(g->off); free(g->ab); free(g->nt); free(g);
}

#ifdef GEN_MAIN
int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s N L SEED OUT.fasta [ab_mode] [orphan_p]\n", argv[0]); return 2; }
  uint64_t n = strtoull(argv[1], 0, 10); uint32_t L = (uint32_t)atoi(argv[2]); uint64_t seed = strtoull(argv[3], 0, 10);
  int mode = argc > 5 ? atoi(argv[5]) : 0; double op = argc > 6 ? atof(argv[6]) : 0.2;
  gen_set *g = gen_create(n, L, seed, mode, op);
  if (!g) return 1;
  int rc = gen_write_fasta(g, argv[4], 0, g->n);
  gen_free(g);
  return rc ? 1 : 0;
}
#endif
