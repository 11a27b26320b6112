/* tools/canon.c — canonical form of a swarm cluster file and its SHA-256 (bench/test infrastructure).
 *
 * "sort ids within each line, then sort lines, LC_ALL=C, and cmp".  Cluster ids may permute
 * between implementations and members may be listed in any order; the canonical text is what must match byte
 * for byte.  canon_sha256() builds that text in memory (ids compared as byte strings, shorter prefix first —
 * the order of `sort` under LC_ALL=C and of Python's sorted() on bytes) and hashes it; blank lines are dropped.
 * sha256_hex() hashes a buffer as it is (for the -s / -i files, which must be byte-identical anyway).
 * Nothing here comes from the reference.  Pure C11.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint32_t h[8];
  uint64_t len;
  unsigned char buf[64];
  uint32_t fill;
} sha256_ctx;

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t ror(uint32_t x, int k) { return (x >> k) | (x << (32 - k)); }

static void sha256_block(sha256_ctx *c, const unsigned char *p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    const uint32_t s0 = ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3);
    const uint32_t s1 = ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = c->h[0], b = c->h[1], cc = c->h[2], d = c->h[3], e = c->h[4], f = c->h[5], g = c->h[6], h = c->h[7];
  for (int i = 0; i < 64; i++) {
    const uint32_t t1 = h + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
    const uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & b) ^ (a & cc) ^ (b & cc));
    h = g; g = f; f = e; e = d + t1; d = cc; cc = b; b = a; a = t1 + t2;
  }
  c->h[0] += a; c->h[1] += b; c->h[2] += cc; c->h[3] += d; c->h[4] += e; c->h[5] += f; c->h[6] += g; c->h[7] += h;
}

static void sha256_init(sha256_ctx *c) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(c->h, iv, sizeof iv);
  c->len = 0;
  c->fill = 0;
}

static void sha256_update(sha256_ctx *c, const void *data, uint64_t n) {
  const unsigned char *p = (const unsigned char *)data;
  c->len += n;
  if (c->fill) {
    while (n && c->fill < 64) { c->buf[c->fill++] = *p++; n--; }
    if (c->fill < 64) return;
    sha256_block(c, c->buf);
    c->fill = 0;
  }
  while (n >= 64) { sha256_block(c, p); p += 64; n -= 64; }
  while (n) { c->buf[c->fill++] = *p++; n--; }
}

static void sha256_final(sha256_ctx *c, char *hex65) {
  const uint64_t bits = c->len * 8;
  unsigned char pad = 0x80;
  sha256_update(c, &pad, 1);
  pad = 0;
  while (c->fill != 56) sha256_update(c, &pad, 1);
  unsigned char lenb[8];
  for (int i = 0; i < 8; i++) lenb[i] = (unsigned char)(bits >> (56 - 8 * i));
  sha256_update(c, lenb, 8);
  for (int i = 0; i < 8; i++) sprintf(hex65 + 8 * i, "%08x", c->h[i]);
  hex65[64] = 0;
}

void sha256_hex(const char *data, uint64_t len, char *hex65) {
  sha256_ctx c;
  sha256_init(&c);
  sha256_update(&c, data, len);
  sha256_final(&c, hex65);
}

typedef struct {
  const char *p;
  uint32_t len;
} tok;

static int tok_cmp(const void *a, const void *b) {
  const tok *x = (const tok *)a, *y = (const tok *)b;
  const uint32_t m = x->len < y->len ? x->len : y->len;
  const int r = memcmp(x->p, y->p, m);
  if (r) return r;
  return x->len < y->len ? -1 : (x->len > y->len ? 1 : 0);
}

typedef struct {
  uint64_t off;
  uint64_t len;
} line_t;

static const char *g_text;
static int line_cmp(const void *a, const void *b) {
  const line_t *x = (const line_t *)a, *y = (const line_t *)b;
  const uint64_t m = x->len < y->len ? x->len : y->len;
  const int r = memcmp(g_text + x->off, g_text + y->off, m);
  if (r) return r;
  return x->len < y->len ? -1 : (x->len > y->len ? 1 : 0);
}

/* canonical text of `text` -> SHA-256 in hex65 (65 bytes); returns the number of non-blank lines, or -1.
 * Tokens are separated by blanks/tabs; lines by \n (a \r before it is dropped). */
int64_t canon_sha256(const char *text, uint64_t len, char *hex65) {
  char *out = (char *)malloc(len + 2);
  uint64_t n_lines = 0, cap_lines = 1 << 16, cap_tok = 1 << 12;
  line_t *lines = (line_t *)malloc(cap_lines * sizeof(line_t));
  tok *toks = (tok *)malloc(cap_tok * sizeof(tok));
  if (!out || !lines || !toks) { free(out); free(lines); free(toks); return -1; }
  uint64_t w = 0, i = 0;
  while (i < len) {
    uint64_t e = i;
    while (e < len && text[e] != '\n') e++;
    uint64_t nt = 0, j = i;
    while (j < e) {
      while (j < e && (text[j] == ' ' || text[j] == '\t' || text[j] == '\r')) j++;
      uint64_t k = j;
      while (k < e && text[k] != ' ' && text[k] != '\t' && text[k] != '\r') k++;
      if (k > j) {
        if (nt == cap_tok) { cap_tok *= 2; toks = (tok *)realloc(toks, cap_tok * sizeof(tok)); if (!toks) { free(out); free(lines); return -1; } }
        toks[nt].p = text + j; toks[nt].len = (uint32_t)(k - j); nt++;
      }
      j = k;
    }
    if (nt) {
      if (nt > 1) qsort(toks, nt, sizeof(tok), tok_cmp);
      if (n_lines == cap_lines) { cap_lines *= 2; lines = (line_t *)realloc(lines, cap_lines * sizeof(line_t)); if (!lines) { free(out); free(toks); return -1; } }
      lines[n_lines].off = w;
      for (uint64_t t = 0; t < nt; t++) {
        if (t) out[w++] = ' ';
        memcpy(out + w, toks[t].p, toks[t].len);
        w += toks[t].len;
      }
      lines[n_lines].len = w - lines[n_lines].off;
      n_lines++;
    }
    i = e + 1;
  }
  g_text = out;
  qsort(lines, n_lines, sizeof(line_t), line_cmp);
  sha256_ctx c;
  sha256_init(&c);
  const char nl = '\n';
  for (uint64_t l = 0; l < n_lines; l++) {
    sha256_update(&c, out + lines[l].off, lines[l].len);
    sha256_update(&c, &nl, 1);
  }
  sha256_final(&c, hex65);
  free(out); free(lines); free(toks);
  return (int64_t)n_lines;
}

#ifdef CANON_MAIN
int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s FILE [raw]\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *b = (char *)malloc((size_t)n + 1);
  if (fread(b, 1, (size_t)n, f) != (size_t)n) return 1;
  fclose(f);
  char hex[65];
  if (argc > 2) sha256_hex(b, (uint64_t)n, hex);
  else if (canon_sha256(b, (uint64_t)n, hex) < 0) return 1;
  printf("%s\n", hex);
  free(b);
  return 0;
}
#endif
