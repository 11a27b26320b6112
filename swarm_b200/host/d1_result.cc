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
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
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
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t s = no_of_seed[swarm_of[i]];
    if (s == 0xFFFFFFFFu) { g_err = "assemble: swarm_of does not point at a seed"; delete r; return 1; }
    r->swarm_no[i] = s;
    r->size[s]++; r->mass[s] += db.abundance[i]; r->sumlen[s] += db.len[i];
    if (db.abundance[i] == 1) r->singletons[s]++;
    r->maxgen[s] = std::max(r->maxgen[s], generation[i]);
  }
  for (uint32_t s = 0; s < ns; ++s) { r->own_size[s] = r->size[s]; r->first[s + 1] = r->first[s] + r->size[s]; }
  // members in list order: counting sort by swarm keeps ids ascending, then stable sort by generation
  r->members.resize(n);
  {
    std::vector<uint64_t> cur(r->first.begin(), r->first.end() - 1);
    for (uint32_t i = 0; i < n; ++i) r->members[cur[r->swarm_no[i]]++] = i;
    for (uint32_t s = 0; s < ns; ++s)
      if (r->maxgen[s] > 1)
        std::stable_sort(r->members.begin() + static_cast<int64_t>(r->first[s]),
                         r->members.begin() + static_cast<int64_t>(r->first[s + 1]),
                         [&](uint32_t a, uint32_t b) { return generation[a] < generation[b]; });
  }
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
  std::string s;
  s.reserve(static_cast<size_t>(r->n) * 16);
  if (mothur) s += "swarm_" + std::to_string(differences) + "\t" + std::to_string(r->swarms_adjusted);
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw) {
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
  if (mothur) s += '\n';
  return give(s, out, out_len);
}

// -s : src/algod1.cc:1043-1062
int swbh_write_stats(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  std::string s;
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw) {
    if (r->attached[sw]) continue;
    s += std::to_string(r->size[sw]) + "\t" + std::to_string(r->mass[sw]) + "\t";
    swb::append_id_noabundance(s, dbh->db, r->seed[sw], o);
    s += "\t" + std::to_string(dbh->db.abundance[r->seed[sw]]) + "\t" + std::to_string(r->singletons[sw]) + "\t" +
         std::to_string(r->maxgen[sw]) + "\t" + std::to_string(r->maxgen[sw]) + "\n";
  }
  return give(s, out, out_len);
}

// -i : src/algod1.cc:990-1040
int swbh_write_structure(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  std::string s;
  uint32_t cluster_no = 0;
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw) {
    if (r->attached[sw]) continue;
    const uint32_t seed = r->seed[sw];
    swb::for_each_member(*r, sw, [&](uint32_t a) {
      if (a == seed) return;
      const uint32_t gp = r->graft_cand[a];
      if (gp != 0xFFFFFFFFu) {
        swb::append_id_noabundance(s, dbh->db, gp, o); s += '\t';
        swb::append_id_noabundance(s, dbh->db, a, o);
        s += "\t2\t" + std::to_string(cluster_no + 1) + "\t" + std::to_string(r->generation[gp] + 1) + "\n";
      }
      const uint32_t par = r->parent[a];
      if (par != 0xFFFFFFFFu) {
        swb::append_id_noabundance(s, dbh->db, par, o); s += '\t';
        swb::append_id_noabundance(s, dbh->db, a, o);
        s += "\t1\t" + std::to_string(cluster_no + 1) + "\t" + std::to_string(r->generation[a]) + "\n";
      }
    });
    ++cluster_no;
  }
  return give(s, out, out_len);
}

// -w : src/algod1.cc:937-987
int swbh_write_seeds(const swbh_db *dbh, const swbh_result *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  const swb::AmpliconDb &db = dbh->db;
  std::vector<uint32_t> sorter(r->seed.size());
  std::iota(sorter.begin(), sorter.end(), 0u);
  std::sort(sorter.begin(), sorter.end(), [&](uint32_t x, uint32_t y) {
    if (r->mass[x] != r->mass[y]) return r->mass[x] > r->mass[y];
    return std::strcmp(db.header(r->seed[x]), db.header(r->seed[y])) < 0;
  });
  std::string s;
  for (uint32_t sw : sorter) {
    if (r->attached[sw]) continue;
    s += '>';
    swb::append_id_new_abundance(s, db, r->seed[sw], r->mass[sw], o);
    s += '\n';
    swb::append_sequence(s, db, r->seed[sw]);
    s += '\n';
  }
  return give(s, out, out_len);
}

// -j : src/algod1.cc:755-788 (rows already sorted ascending by the engine)
int swbh_write_network(const swbh_db *dbh, const uint64_t *row_ptr, const uint32_t *col, int usearch, int64_t append,
                       char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append;
  std::string s;
  for (uint32_t i = 0; i < dbh->db.n; ++i)
    for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
      swb::append_id(s, dbh->db, i, o); s += '\t';
      swb::append_id(s, dbh->db, col[k], o); s += '\n';
    }
  return give(s, out, out_len);
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
  std::string s;
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw) {
    s += std::to_string(r->size[sw]) + "\t" + std::to_string(r->mass[sw]) + "\t";
    swb::append_id_noabundance(s, dbh->db, r->seed[sw], o);
    s += "\t" + std::to_string(dbh->db.abundance[r->seed[sw]]) + "\t" + std::to_string(r->singletons[sw]) + "\t" +
         std::to_string(std::max<uint32_t>(1, r->maxgen[sw])) + "\t" + std::to_string(r->maxradius[sw]) + "\n";
  }
  return give(s, out, out_len);
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
  std::string s;
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw)
    for (uint64_t k = 0; k < r->own_size[sw]; ++k) {
      const uint32_t par = r->members[r->first[sw] + k];
      for (uint64_t c = cstart[par]; c < cstart[par + 1]; ++c) {
        const uint32_t a = child[c];
        swb::append_id_noabundance(s, dbh->db, par, o); s += '\t';
        swb::append_id_noabundance(s, dbh->db, a, o);
        s += "\t" + std::to_string(r->pdiff[a]) + "\t" + std::to_string(sw + 1) + "\t" + std::to_string(r->generation[a]) + "\n";
      }
    }
  return give(s, out, out_len);
}

}  // extern "C"
