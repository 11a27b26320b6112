// swarm_b200/host/derep.cc — d = 0: turn the engine's dereplication arrays into swarm's outputs.
//
// swb200_d0_dereplicate() returns rep[i] (the first amplicon with i's sequence) and the cluster sums at the
// representatives.  This file orders the clusters the way the reference does (mass descending, then seed index:
// /root/reference src/derep.cc:74-98 `sort_seeds`), rebuilds the member chains (`nextseqtab`: members of a cluster in
// index order, :322-326) and mirrors the d = 0 writers byte for byte: swarms :244-273 / mothur :206-241, seeds :190-203,
// UCLUST-like :145-187, structure :121-142, statistics :103-118.  No alignment is needed: all members are identical.
#include "../../include/swarm_b200_host.h"
#include "amplicon_db.h"
#include "result.h"

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

namespace swb { int give_text(const std::string &s, char **out, uint64_t *out_len); void set_host_error(const std::string &e); }

struct swbh_derep {
  std::vector<uint32_t> seed;           // clusters in output order
  std::vector<uint64_t> mass;
  std::vector<uint32_t> size, singletons;
  std::vector<uint64_t> first;          // clusters + 1 offsets into members
  std::vector<uint32_t> members;        // per cluster: seed, then the other members, ids ascending
  uint64_t heaviest = 0;
  uint32_t largest = 0;
};

extern "C" {

int swbh_d0_assemble(const swbh_db *dbh, const uint32_t *rep, const uint64_t *mass, const uint32_t *size,
                     const uint32_t *singletons, swbh_derep **out) {
  const uint32_t n = dbh->db.n;
  auto *r = new swbh_derep();
  for (uint32_t i = 0; i < n; ++i) {
    if (rep[i] > i || rep[rep[i]] != rep[i]) { swb::set_host_error("d0 assemble: rep is not a first-occurrence map"); delete r; return 1; }
    if (rep[i] == i) r->seed.push_back(i);
  }
  std::sort(r->seed.begin(), r->seed.end(), [&](uint32_t a, uint32_t b) { return mass[a] != mass[b] ? mass[a] > mass[b] : a < b; });
  const size_t k = r->seed.size();
  std::vector<uint32_t> rank(n, 0);
  r->mass.resize(k); r->size.resize(k); r->singletons.resize(k); r->first.assign(k + 1, 0);
  for (size_t c = 0; c < k; ++c) {
    const uint32_t s = r->seed[c];
    rank[s] = static_cast<uint32_t>(c);
    r->mass[c] = mass[s]; r->size[c] = size[s]; r->singletons[c] = singletons[s];
    r->first[c + 1] = r->first[c] + size[s];
    r->heaviest = std::max(r->heaviest, mass[s]);
    r->largest = std::max(r->largest, size[s]);
  }
  if (r->first[k] != n) { swb::set_host_error("d0 assemble: cluster sizes do not add up"); delete r; return 1; }
  r->members.resize(n);
  std::vector<uint64_t> cur(r->first.begin(), r->first.end() - 1);
  for (uint32_t i = 0; i < n; ++i) r->members[cur[rank[rep[i]]]++] = i;      // ascending ids inside a cluster
  *out = r;
  return 0;
}

void swbh_derep_free(swbh_derep *r) { delete r; }
uint64_t swbh_derep_clusters(const swbh_derep *r) { return r->seed.size(); }
uint32_t swbh_derep_largest(const swbh_derep *r) { return r->largest; }
uint64_t swbh_derep_heaviest(const swbh_derep *r) { return r->heaviest; }

// The writers format clusters [c0, c1) per worker and concatenate in order (result.h: parallel_text): identical bytes for any
// worker count; 10 M reads are hundreds of MB of text.
// -o / -r : src/derep.cc:244-273, :206-241
int swbh_d0_write_swarms(const swbh_db *dbh, const swbh_derep *r, int mothur, int usearch, int64_t append, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append;
  std::string head, tail;
  if (mothur) { head = "swarm_0\t" + std::to_string(r->seed.size()); tail = "\n"; }
  return swb::parallel_text(
      r->seed.size(), [&](uint64_t c) { return r->first[c]; },
      [&](uint64_t c0, uint64_t c1, std::string &s) {
        for (uint64_t c = c0; c < c1; ++c) {
          for (uint64_t m = r->first[c]; m < r->first[c + 1]; ++m) {
            if (mothur) s += m == r->first[c] ? '\t' : ',';
            else if (m != r->first[c]) s += ' ';
            swb::append_id(s, dbh->db, r->members[m], o);
          }
          if (!mothur) s += '\n';
        }
      },
      head, tail, out, out_len);
}

// -w : src/derep.cc:190-203 (cluster order, not re-sorted)
int swbh_d0_write_seeds(const swbh_db *dbh, const swbh_derep *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  return swb::parallel_text(
      r->seed.size(), [](uint64_t c) { return c * 16; },
      [&](uint64_t c0, uint64_t c1, std::string &s) {
        for (uint64_t c = c0; c < c1; ++c) {
          s += '>';
          swb::append_id_new_abundance(s, dbh->db, r->seed[c], r->mass[c], o);
          s += '\n';
          swb::append_sequence(s, dbh->db, r->seed[c]);
          s += '\n';
        }
      },
      "", "", out, out_len);
}

// -u : src/derep.cc:145-187
int swbh_d0_write_uclust(const swbh_db *dbh, const swbh_derep *r, int usearch, int64_t append, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append;
  const swb::AmpliconDb &db = dbh->db;
  return swb::parallel_text(
      r->seed.size(), [&](uint64_t c) { return r->first[c] * 4 + c * 8; },
      [&](uint64_t c0, uint64_t c1, std::string &s) {
        for (uint64_t c = c0; c < c1; ++c) {
          const uint32_t seed = r->seed[c];
          s += "C\t"; swb::append_uint(s, c); s += '\t'; swb::append_uint(s, r->size[c]); s += "\t*\t*\t*\t*\t*\t";
          swb::append_id(s, db, seed, o);
          s += "\t*\nS\t"; swb::append_uint(s, c); s += '\t'; swb::append_uint(s, db.len[seed]); s += "\t*\t*\t*\t*\t*\t";
          swb::append_id(s, db, seed, o);
          s += "\t*\n";
          for (uint64_t m = r->first[c] + 1; m < r->first[c + 1]; ++m) {
            const uint32_t a = r->members[m];
            s += "H\t"; swb::append_uint(s, c); s += '\t'; swb::append_uint(s, db.len[a]); s += "\t100.0\t+\t0\t0\t=\t";
            swb::append_id(s, db, a, o);
            s += '\t';
            swb::append_id(s, db, seed, o);
            s += '\n';
          }
        }
      },
      "", "", out, out_len);
}

// -i : src/derep.cc:121-142
int swbh_d0_write_structure(const swbh_db *dbh, const swbh_derep *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  return swb::parallel_text(
      r->seed.size(), [&](uint64_t c) { return (r->first[c] - c) * 3; },      // one line per member that is not a seed
      [&](uint64_t c0, uint64_t c1, std::string &s) {
        for (uint64_t c = c0; c < c1; ++c)
          for (uint64_t m = r->first[c] + 1; m < r->first[c + 1]; ++m) {
            swb::append_id_noabundance(s, dbh->db, r->seed[c], o); s += '\t';
            swb::append_id_noabundance(s, dbh->db, r->members[m], o);
            s += "\t0\t"; swb::append_uint(s, c + 1); s += "\t0\n";
          }
      },
      "", "", out, out_len);
}

// -s : src/derep.cc:103-118
int swbh_d0_write_stats(const swbh_db *dbh, const swbh_derep *r, int usearch, char **out, uint64_t *out_len) {
  swb::DbOptions o; o.usearch_abundance = usearch != 0;
  return swb::parallel_text(
      r->seed.size(), [](uint64_t c) { return c * 4; },
      [&](uint64_t c0, uint64_t c1, std::string &s) {
        for (uint64_t c = c0; c < c1; ++c) {
          swb::append_uint(s, r->size[c]); s += '\t'; swb::append_uint(s, r->mass[c]); s += '\t';
          swb::append_id_noabundance(s, dbh->db, r->seed[c], o);
          s += '\t'; swb::append_uint(s, dbh->db.abundance[r->seed[c]]); s += '\t'; swb::append_uint(s, r->singletons[c]); s += "\t0\t0\n";
        }
      },
      "", "", out, out_len);
}

}  // extern "C"
