// swarm_b200/host/d1_result.cc — turn the engine's per-amplicon arrays into swarm's d=1 outputs.
//
// The engine returns swarm_of / generation / parent (and graft_cand for --fastidious).  This file
// rebuilds what the reference keeps in `ampinfo_s.next` + `swarminfo_s` (/root/reference
// src/algod1.cc:87-116) and mirrors its writers byte for byte (src/algod1.cc:755-1062).
//   * swarms are numbered by seed index (the greedy loop starts swarms in index order, :1185-1192);
//   * inside a swarm the list order is: seed, then generation by generation, each generation sorted
//     by amplicon id (:1215-1250);
//   * grafting: pairs (graft_cand[l], l) sorted by (parent, child); a light swarm is attached once, to
//     the swarm of its first pair, at the tail of that swarm's list; later pairs of the same light swarm
//     clear graft_cand (:274-336, attach :214-241).
#include "../../include/swarm_b200_host.h"
#include "amplicon_db.h"
#include "result.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_err;

int give(const std::string &s, char **out, uint64_t *out_len) {
  char *p = static_cast<char *>(std::malloc(s.size() + 1));
  if (!p) { g_err = "out of memory"; return 1; }
  std::memcpy(p, s.data(), s.size());
  p[s.size()] = '\0';
  *out = p;
  if (out_len) *out_len = s.size();
  return 0;
}

}  // namespace

namespace swb {
int give_text(const std::string &s, char **out, uint64_t *out_len) { return give(s, out, out_len); }
void set_host_error(const std::string &e) { g_err = e; }

namespace { std::atomic<uint64_t> g_writer_grain{200000}; }
void set_writer_grain(uint64_t weight) { g_writer_grain.store(weight ? weight : 1); }

void append_uint(std::string &s, uint64_t v) {
  char buf[24];
  int k = 24;
  do { buf[--k] = static_cast<char>('0' + v % 10); v /= 10; } while (v);
  s.append(buf + k, static_cast<size_t>(24 - k));
}

int parallel_text(uint64_t n_units, const std::function<uint64_t(uint64_t)> &weight_before,
                  const std::function<void(uint64_t, uint64_t, std::string &)> &body, const std::string &head, const std::string &tail,
                  char **out, uint64_t *out_len) {
  const uint64_t total = n_units ? weight_before(n_units) : 0;
  unsigned T = host_threads();
  if (total < g_writer_grain.load() || n_units < 2) T = 1;
  T = static_cast<unsigned>(std::min<uint64_t>(T, n_units ? n_units : 1));
  // range t = units [cut[t], cut[t+1]): the first unit whose weight_before reaches t / T of the total (binary search)
  std::vector<uint64_t> cut(T + 1, n_units);
  cut[0] = 0;
  for (unsigned t = 1; t < T; ++t) {
    const uint64_t want = total / T * t;
    uint64_t lo = cut[t - 1], hi = n_units;
    while (lo < hi) { const uint64_t mid = (lo + hi) / 2; if (weight_before(mid) < want) lo = mid + 1; else hi = mid; }
    cut[t] = lo;
  }
  std::vector<std::string> part(T);
  auto work = [&](unsigned t) {
    if (cut[t] < cut[t + 1]) {
      part[t].reserve(static_cast<size_t>((weight_before(cut[t + 1]) - weight_before(cut[t])) * 12 + 64));
      body(cut[t], cut[t + 1], part[t]);
    }
  };
  if (T <= 1) work(0);
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  }
  uint64_t bytes = head.size() + tail.size();
  std::vector<uint64_t> at(T + 1, head.size());
  for (unsigned t = 0; t < T; ++t) { at[t + 1] = at[t] + part[t].size(); bytes += part[t].size(); }
  char *p = static_cast<char *>(std::malloc(bytes + 1));
  if (!p) { g_err = "out of memory"; return 1; }
  std::memcpy(p, head.data(), head.size());
  auto copy = [&](unsigned t) { std::memcpy(p + at[t], part[t].data(), part[t].size()); };
  if (T <= 1) copy(0);
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(copy, t);
    for (auto &th : pool) th.join();
  }
  std::memcpy(p + at[T], tail.data(), tail.size());
  p[bytes] = '\0';
  *out = p;
  if (out_len) *out_len = bytes;
  return 0;
}
}  // namespace swb

extern "C" {

const char *swbh_last_error(void) { return g_err.c_str(); }

int swbh_db_read_fasta(const char *path, int usearch, int64_t append, int check_dup, swbh_db **out) {
  auto *h = new swbh_db();
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append; o.check_duplicate_sequences = check_dup != 0;
  const std::string e = swb::db_read_file(path ? path : "-", o, h->db);
  if (!e.empty()) { g_err = e; delete h; *out = nullptr; return 1; }
  *out = h;
  return 0;
}
int swbh_db_parse(const char *text, uint64_t size, int usearch, int64_t append, int check_dup, swbh_db **out) {
  auto *h = new swbh_db();
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append; o.check_duplicate_sequences = check_dup != 0;
  const std::string e = swb::db_parse(text, size, o, h->db);
  if (!e.empty()) { g_err = e; delete h; *out = nullptr; return 1; }
  *out = h;
  return 0;
}
void swbh_set_threads(int threads) { swb::set_ingest_threads(threads); }
void swbh_db_free(swbh_db *db) { delete db; }
uint32_t swbh_db_count(const swbh_db *d) { return d->db.n; }
uint32_t swbh_db_longest(const swbh_db *d) { return d->db.longest; }
uint32_t swbh_db_stride_words(const swbh_db *d) { return d->db.stride; }
uint64_t swbh_db_nucleotides(const swbh_db *d) { return d->db.nucleotides; }
const uint64_t *swbh_db_words(const swbh_db *d) { return d->db.words.data(); }
const uint32_t *swbh_db_lengths(const swbh_db *d) { return d->db.len.data(); }
const uint64_t *swbh_db_abundances(const swbh_db *d) { return d->db.abundance.data(); }
const char *swbh_db_header(const swbh_db *d, uint32_t i) { return d->db.header(i); }
void swbh_free(void *p) { std::free(p); }

// the arguments of swb200_load_db_compact(): 16-bit lengths and abundance runs.  Returns the number of runs, or 0 when
// the database cannot be expressed that way (a sequence longer than 65 535 nt) — the caller then uses swb200_load_db.
uint32_t swbh_db_compact(swbh_db *d, const uint16_t **len16, const uint64_t **run_abundance, const uint32_t **run_start) {
  const swb::AmpliconDb &db = d->db;
  if (db.n == 0 || db.longest > 65535) return 0;
  if (d->len16.size() != db.n) {
    d->len16.resize(db.n);
    d->run_abundance.clear(); d->run_start.clear();
    for (uint32_t i = 0; i < db.n; ++i) {
      d->len16[i] = static_cast<uint16_t>(db.len[i]);
      if (i == 0 || db.abundance[i] != db.abundance[i - 1]) { d->run_abundance.push_back(db.abundance[i]); d->run_start.push_back(i); }
    }
    d->run_start.push_back(db.n);
  }
  *len16 = d->len16.data(); *run_abundance = d->run_abundance.data(); *run_start = d->run_start.data();
  return static_cast<uint32_t>(d->run_abundance.size());
}

int swbh_d1_assemble(const swbh_db *dbh, const uint32_t *swarm_of, const uint32_t *generation, const uint32_t *parent,
                     const uint32_t *graft_cand, uint64_t boundary, swbh_result **out) {
  (void)boundary;
  const swb::AmpliconDb &db = dbh->db;
  const uint32_t n = db.n;
  auto *r = new swbh_result();
  r->n = n;
  r->swarm_no.assign(n, 0);
  r->generation.assign(generation, generation + n);
  r->parent.assign(parent, parent + n);
  r->graft_cand.assign(n, 0xFFFFFFFFu);
  // number swarms by seed index
  std::vector<uint32_t> no_of_seed(n, 0xFFFFFFFFu);
  uint32_t ns = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (swarm_of[i] >= n) { g_err = "assemble: swarm_of out of range"; delete r; return 1; }
    if (swarm_of[i] == i) { no_of_seed[i] = ns++; r->seed.push_back(i); }
  }
  r->size.assign(ns, 0); r->singletons.assign(ns, 0); r->maxgen.assign(ns, 0);
  r->mass.assign(ns, 0); r->sumlen.assign(ns, 0); r->attached.assign(ns, 0);
  r->own_size.assign(ns, 0); r->first.assign(static_cast<size_t>(ns) + 1, 0);
  r->grafted.assign(ns, {});
  // Both passes below run on several workers for large results (10 M amplicons: 0.85 s on one thread).  Pass 1: worker t owns the
  // SWARMS [s_t, s_t+1) — it scans all amplicons (a sequential read) and accumulates only its own swarms' sums, so no two workers
  // touch the same counter.  Pass 2 (members in list order) uses the same ownership: the counting sort by swarm keeps ids
  // ascending, then every swarm with more than one generation is stably sorted by generation.
  const unsigned T = (n >= swb::g_writer_grain.load() && ns >= 2) ? std::max(1u, std::min({swb::host_threads(), 16u, ns})) : 1u;
  auto swarm_range = [&](unsigned t, uint32_t &s0, uint32_t &s1) {
    s0 = static_cast<uint32_t>(static_cast<uint64_t>(ns) * t / T);
    s1 = static_cast<uint32_t>(static_cast<uint64_t>(ns) * (t + 1) / T);
  };
  std::vector<int> bad(T, 0);
  auto run = [&](const std::function<void(unsigned)> &f) {
    if (T <= 1) { f(0); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(f, t);
    for (auto &th : pool) th.join();
  };
  {                                       // swarm number of every amplicon (amplicon ranges: disjoint writes)
    run([&](unsigned t) {
      const uint32_t i0 = static_cast<uint32_t>(static_cast<uint64_t>(n) * t / T), i1 = static_cast<uint32_t>(static_cast<uint64_t>(n) * (t + 1) / T);
      for (uint32_t i = i0; i < i1; ++i) {
        const uint32_t s = no_of_seed[swarm_of[i]];
        if (s == 0xFFFFFFFFu) { bad[t] = 1; return; }
        r->swarm_no[i] = s;
      }
    });
    for (unsigned t = 0; t < T; ++t)
      if (bad[t]) { g_err = "assemble: swarm_of does not point at a seed"; delete r; return 1; }
  }
  run([&](unsigned t) {
    uint32_t s0, s1;
    swarm_range(t, s0, s1);
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t s = r->swarm_no[i];
      if (s < s0 || s >= s1) continue;
      r->size[s]++; r->mass[s] += db.abundance[i]; r->sumlen[s] += db.len[i];
      if (db.abundance[i] == 1) r->singletons[s]++;
      r->maxgen[s] = std::max(r->maxgen[s], generation[i]);
    }
  });
  for (uint32_t s = 0; s < ns; ++s) { r->own_size[s] = r->size[s]; r->first[s + 1] = r->first[s] + r->size[s]; }
  r->members.resize(n);
  run([&](unsigned t) {
    uint32_t s0, s1;
    swarm_range(t, s0, s1);
    if (s0 == s1) return;
    std::vector<uint64_t> cur(r->first.begin() + s0, r->first.begin() + s1);
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t s = r->swarm_no[i];
      if (s >= s0 && s < s1) r->members[cur[s - s0]++] = i;
    }
    for (uint32_t s = s0; s < s1; ++s)
      if (r->maxgen[s] > 1)
        std::stable_sort(r->members.begin() + static_cast<int64_t>(r->first[s]),
                         r->members.begin() + static_cast<int64_t>(r->first[s + 1]),
                         [&](uint32_t a, uint32_t b) { return generation[a] < generation[b]; });
  });
  for (uint32_t s = 0; s < ns; ++s) {
    r->largest = std::max(r->largest, r->size[s]);
    r->maxgen_all = std::max(r->maxgen_all, r->maxgen[s]);
  }
  r->swarms_adjusted = ns;
  // grafting (src/algod1.cc:274-336)
  if (graft_cand != nullptr) {
    struct Pair { uint32_t parent, child; };
    std::vector<Pair> pairs;
    for (uint32_t i = 0; i < n; ++i)
      if (graft_cand[i] != 0xFFFFFFFFu) { pairs.push_back({graft_cand[i], i}); r->graft_cand[i] = graft_cand[i]; }
    std::sort(pairs.begin(), pairs.end(), [](const Pair &a, const Pair &b) {
      return a.parent != b.parent ? a.parent < b.parent : a.child < b.child;
    });
    for (const Pair &p : pairs) {
      const uint32_t ls = r->swarm_no[p.child], hs = r->swarm_no[p.parent];
      if (r->attached[ls]) { r->graft_cand[p.child] = 0xFFFFFFFFu; continue; }
      r->grafted[hs].push_back(ls);
      r->size[hs] += r->size[ls]; r->singletons[hs] += r->singletons[ls];
      r->mass[hs] += r->mass[ls]; r->sumlen[hs] += r->sumlen[ls];
      r->attached[ls] = 1;
      r->largest = std::max(r->largest, r->size[hs]);
      r->swarms_adjusted--;
      r->grafts++;
    }
  }
  *out = r;
  return 0;
}

void swbh_result_free(swbh_result *r) { delete r; }
uint64_t swbh_result_swarms(const swbh_result *r) { return r->swarms_adjusted; }
uint32_t swbh_result_largest(const swbh_result *r) { return r->largest; }
uint32_t swbh_result_maxgen(const swbh_result *r) { return r->maxgen_all; }
uint64_t swbh_result_grafts(const swbh_result *r) { return r->grafts; }

// -o / -r : src/algod1.cc:791-849
int swbh_write_swarms(const swbh_db *dbh, const swbh_result *r, int mothur, int64_t differences, int usearch,
                      int64_t append, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append;
  std::string head, tail;
  if (mothur) { head = "swarm_" + std::to_string(differences) + "\t" + std::to_string(r->swarms_adjusted); tail = "\n"; }
  return swb::parallel_text(
      r->seed.size(), [&](uint64_t sw) { return r->first[sw]; },
      [&](uint64_t s0, uint64_t s1, std::string &s) {
        for (uint32_t sw = static_cast<uint32_t>(s0); sw < s1; ++sw) {
          if (r->attached[sw]) continue;
          bool first = true;
          swb::for_each_member(*r, sw, [&](uint32_t a) {
            if (mothur) s += first ? '\t' : ',';
            else if (!first) s += ' ';
            first = false;
            swb::append_id(s, dbh->db, a, o);
          });
          if (!mothur) s += '\n';
        }
      },
      head, tail, out, out_len);
}

// -s : src/algod1.cc:1043-1062
int swbh_write_stats(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  return swb::parallel_text(
      r->seed.size(), [](uint64_t sw) { return sw * 4; },
      [&](uint64_t s0, uint64_t s1, std::string &s) {
        for (uint32_t sw = static_cast<uint32_t>(s0); sw < s1; ++sw) {
          if (r->attached[sw]) continue;
          swb::append_uint(s, r->size[sw]); s += '\t'; swb::append_uint(s, r->mass[sw]); s += '\t';
          swb::append_id_noabundance(s, dbh->db, r->seed[sw], o);
          s += '\t'; swb::append_uint(s, dbh->db.abundance[r->seed[sw]]); s += '\t'; swb::append_uint(s, r->singletons[sw]);
          s += '\t'; swb::append_uint(s, r->maxgen[sw]); s += '\t'; swb::append_uint(s, r->maxgen[sw]); s += '\n';
        }
      },
      "", "", out, out_len);
}

// -i : src/algod1.cc:990-1040
int swbh_write_structure(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  // output number of a swarm = swarms not attached to another one before it (+ 1)
  std::vector<uint32_t> number(r->seed.size() + 1, 0);
  for (size_t sw = 0; sw < r->seed.size(); ++sw) number[sw + 1] = number[sw] + (r->attached[sw] ? 0u : 1u);
  return swb::parallel_text(
      r->seed.size(), [&](uint64_t sw) { return (r->first[sw] - sw) * 3; },      // one line per member that is not a seed
      [&](uint64_t s0, uint64_t s1, std::string &s) {
        for (uint32_t sw = static_cast<uint32_t>(s0); sw < s1; ++sw) {
          if (r->attached[sw]) continue;
          const uint32_t seed = r->seed[sw], cluster_no = number[sw];
          swb::for_each_member(*r, sw, [&](uint32_t a) {
            if (a == seed) return;
            const uint32_t gp = r->graft_cand[a];
            if (gp != 0xFFFFFFFFu) {
              swb::append_id_noabundance(s, dbh->db, gp, o); s += '\t';
              swb::append_id_noabundance(s, dbh->db, a, o);
              s += "\t2\t"; swb::append_uint(s, cluster_no + 1); s += '\t'; swb::append_uint(s, r->generation[gp] + 1); s += '\n';
            }
            const uint32_t par = r->parent[a];
            if (par != 0xFFFFFFFFu) {
              swb::append_id_noabundance(s, dbh->db, par, o); s += '\t';
              swb::append_id_noabundance(s, dbh->db, a, o);
              s += "\t1\t"; swb::append_uint(s, cluster_no + 1); s += '\t'; swb::append_uint(s, r->generation[a]); s += '\n';
            }
          });
        }
      },
      "", "", out, out_len);
}

// -w : src/algod1.cc:937-987
int swbh_write_seeds(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  const swb::AmpliconDb &db = dbh->db;
  // order: mass descending, then header ascending (strcmp).  The first 8 header bytes, big-endian, decide almost every comparison
  // without touching the header text (2.4 M seeds at 10 M amplicons: the strcmp-only sort took seconds)
  struct Key { uint64_t mass, head; uint32_t sw; };
  std::vector<Key> keys(r->seed.size());
  for (uint32_t sw = 0; sw < keys.size(); ++sw) {
    const char *h = db.header(r->seed[sw]);
    uint64_t v = 0;
    int k = 0;
    for (; k < 8 && h[k]; ++k) v = (v << 8) | static_cast<unsigned char>(h[k]);
    v <<= 8 * (8 - k);                                        // a shorter header sorts first, as with strcmp
    keys[sw] = {r->mass[sw], v, sw};
  }
  auto less = [&](const Key &x, const Key &y) {
    if (x.mass != y.mass) return x.mass > y.mass;
    if (x.head != y.head) return x.head < y.head;
    return std::strcmp(db.header(r->seed[x.sw]), db.header(r->seed[y.sw])) < 0;
  };
  {
    // T sorted runs on T workers, then pairwise merges (log2 T passes, the merges of a pass in parallel); headers are unique, so
    // the order is total and the result does not depend on T
    unsigned T = keys.size() >= swb::g_writer_grain.load() ? std::max(1u, std::min(swb::host_threads(), 16u)) : 1u;
    while (T & (T - 1)) T &= T - 1;                          // a power of two
    std::vector<size_t> cut(T + 1);
    for (unsigned t = 0; t <= T; ++t) cut[t] = keys.size() * t / T;
    auto each = [&](unsigned count, const std::function<void(unsigned)> &f) {
      if (count <= 1) { if (count) f(0); return; }
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < count; ++t) pool.emplace_back(f, t);
      for (auto &th : pool) th.join();
    };
    each(T, [&](unsigned t) { std::sort(keys.begin() + static_cast<int64_t>(cut[t]), keys.begin() + static_cast<int64_t>(cut[t + 1]), less); });
    for (unsigned width = 1; width < T; width *= 2)
      each(T / (2 * width), [&](unsigned p) {
        const size_t a = cut[2 * width * p], m = cut[2 * width * p + width], b = cut[2 * width * (p + 1)];
        std::inplace_merge(keys.begin() + static_cast<int64_t>(a), keys.begin() + static_cast<int64_t>(m), keys.begin() + static_cast<int64_t>(b), less);
      });
  }
  std::vector<uint32_t> sorter(keys.size());
  for (size_t k = 0; k < keys.size(); ++k) sorter[k] = keys[k].sw;
  return swb::parallel_text(
      sorter.size(), [](uint64_t k) { return k * 16; },
      [&](uint64_t k0, uint64_t k1, std::string &s) {
        for (uint64_t k = k0; k < k1; ++k) {
          const uint32_t sw = sorter[k];
          if (r->attached[sw]) continue;
          s += '>';
          swb::append_id_new_abundance(s, db, r->seed[sw], r->mass[sw], o);
          s += '\n';
          swb::append_sequence(s, db, r->seed[sw]);
          s += '\n';
        }
      },
      "", "", out, out_len);
}

// -j : src/algod1.cc:755-788 (rows already sorted ascending by the engine)
int swbh_write_network(const swbh_db *dbh, const uint64_t *row_ptr, const uint32_t *col, int usearch, int64_t append,
                       char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append;
  return swb::parallel_text(
      dbh->db.n, [&](uint64_t i) { return row_ptr[i] * 2; },
      [&](uint64_t i0, uint64_t i1, std::string &s) {
        for (uint32_t i = static_cast<uint32_t>(i0); i < i1; ++i)
          for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
            swb::append_id(s, dbh->db, i, o); s += '\t';
            swb::append_id(s, dbh->db, col[k], o); s += '\n';
          }
      },
      "", "", out, out_len);
}


// ---- d>1 -----------------------------------------------------------------------------------------
void swbh_scoring(int64_t m, int64_t p, int64_t g, int64_t e, int64_t pen[3]) {   // src/swarm.cc:466-483
  int64_t mis = 2 * m + 2 * p, go = 2 * g, ge = m + 2 * e;
  auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
  const int64_t f = gcd(gcd(mis, go), ge);
  pen[0] = mis / f; pen[1] = go / f; pen[2] = ge / f;
}

int swbh_dn_assemble(const swbh_db *dbh, const uint32_t *swarm_of, const uint32_t *generation, const uint32_t *parent,
                     const uint32_t *pdiff, swbh_result **out) {
  swbh_result *r = nullptr;
  if (swbh_d1_assemble(dbh, swarm_of, generation, parent, nullptr, 0, &r) != 0) return 1;
  const uint32_t n = r->n;
  r->pdiff.assign(pdiff, pdiff + n);
  r->radius.assign(n, 0);
  r->maxradius.assign(r->seed.size(), 0);
  // members are stored generation by generation, so a parent's radius is final before its children's
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw)
    for (uint64_t k = 0; k < r->own_size[sw]; ++k) {
      const uint32_t a = r->members[r->first[sw] + k];
      if (parent[a] != 0xFFFFFFFFu) r->radius[a] = r->radius[parent[a]] + pdiff[a];   // src/algo.cc:493,569
      r->maxradius[sw] = std::max(r->maxradius[sw], r->radius[a]);
    }
  *out = r;
  return 0;
}

// -s at d>1: src/algo.cc:660-674 — maxgen starts at 1 (:401), last column = max radius
int swbh_dn_write_stats(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  return swb::parallel_text(
      r->seed.size(), [](uint64_t sw) { return sw * 4; },
      [&](uint64_t s0, uint64_t s1, std::string &s) {
        for (uint32_t sw = static_cast<uint32_t>(s0); sw < s1; ++sw) {
          swb::append_uint(s, r->size[sw]); s += '\t'; swb::append_uint(s, r->mass[sw]); s += '\t';
          swb::append_id_noabundance(s, dbh->db, r->seed[sw], o);
          s += '\t'; swb::append_uint(s, dbh->db.abundance[r->seed[sw]]); s += '\t'; swb::append_uint(s, r->singletons[sw]);
          s += '\t'; swb::append_uint(s, std::max<uint32_t>(1, r->maxgen[sw])); s += '\t'; swb::append_uint(s, r->maxradius[sw]); s += '\n';
        }
      },
      "", "", out, out_len);
}

// -i at d>1: src/algo.cc:470-484, :573-586 — one line per accepted link, in discovery order: swarm by
// swarm, (sub)seeds in list order, each one's hits in pool order (ascending id)
int swbh_dn_write_structure(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  const uint32_t n = r->n;
  // children of every amplicon, ascending id
  std::vector<uint64_t> cstart(static_cast<size_t>(n) + 1, 0);
  for (uint32_t a = 0; a < n; ++a) if (r->parent[a] != 0xFFFFFFFFu) cstart[r->parent[a] + 1]++;
  for (uint32_t a = 0; a < n; ++a) cstart[a + 1] += cstart[a];
  std::vector<uint32_t> child(cstart[n]);
  { std::vector<uint64_t> cur(cstart.begin(), cstart.end() - 1);
    for (uint32_t a = 0; a < n; ++a) if (r->parent[a] != 0xFFFFFFFFu) child[cur[r->parent[a]]++] = a; }
  return swb::parallel_text(
      r->seed.size(), [&](uint64_t sw) { return (r->first[sw] - sw) * 3; },      // one line per member that is not a seed
      [&](uint64_t s0, uint64_t s1, std::string &s) {
        for (uint32_t sw = static_cast<uint32_t>(s0); sw < s1; ++sw)
          for (uint64_t k = 0; k < r->own_size[sw]; ++k) {
            const uint32_t par = r->members[r->first[sw] + k];
            for (uint64_t c = cstart[par]; c < cstart[par + 1]; ++c) {
              const uint32_t a = child[c];
              swb::append_id_noabundance(s, dbh->db, par, o); s += '\t';
              swb::append_id_noabundance(s, dbh->db, a, o);
              s += '\t'; swb::append_uint(s, r->pdiff[a]); s += '\t'; swb::append_uint(s, sw + 1); s += '\t';
              swb::append_uint(s, r->generation[a]); s += '\n';
            }
          }
      },
      "", "", out, out_len);
}

void swbh_set_writer_grain(uint64_t weight) { swb::set_writer_grain(weight); }

}  // extern "C"
