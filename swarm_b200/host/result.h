// swarm_b200/host/result.h — the host layer's internal view of a database handle and of an assembled
// clustering result (shared by the writers in d1_result.cc, uclust.cc and derep.cc; not part of the C ABI).
#pragma once
#include "amplicon_db.h"

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

struct swbh_db {
  swb::AmpliconDb db;
  // the compact form swb200_load_db_compact() takes, built on first use
  std::vector<uint16_t> len16;
  std::vector<uint64_t> run_abundance;
  std::vector<uint32_t> run_start;
};

struct swbh_result {
  uint32_t n = 0;
  std::vector<uint32_t> swarm_no;       // per amplicon: swarm number (by seed order)
  std::vector<uint32_t> generation, parent, graft_cand, pdiff, radius;
  std::vector<uint32_t> maxradius;       // per swarm (d>1)
  // per swarm (before grafting numbering)
  std::vector<uint32_t> seed, size, singletons, maxgen;
  std::vector<uint64_t> mass, sumlen;
  std::vector<uint8_t> attached;
  std::vector<uint64_t> first;          // offset of the swarm's own members in `members`
  std::vector<uint32_t> own_size;
  std::vector<uint32_t> members;        // own members of every swarm, list order
  std::vector<std::vector<uint32_t>> grafted;   // per heavy swarm: attached light swarm numbers, in attach order
  uint64_t swarms_adjusted = 0, grafts = 0;
  uint32_t largest = 0, maxgen_all = 0;
};

namespace swb {
// Output text built by several workers (the reference writes serially; 10 M amplicons are 117 MB of -o / -s text, 2.4 s of the
// drop-in's 4.8 s wall clock when formatted by one thread): the units [0, n_units) — swarms, rows — are cut into contiguous ranges
// of equal weight (`weight_before(u)` = non-decreasing work before unit u), every worker formats its range into its own buffer with
// `body(u0, u1, text)`, and the buffers are copied, in order, into one malloc'd block behind `head`.  Identical bytes for any
// worker count.  Small outputs (total weight below the grain) stay on the calling thread.
int parallel_text(uint64_t n_units, const std::function<uint64_t(uint64_t)> &weight_before,
                  const std::function<void(uint64_t, uint64_t, std::string &)> &body, const std::string &head, const std::string &tail,
                  char **out, uint64_t *out_len);
void set_writer_grain(uint64_t weight);          // test hook: parallel formatting from this total weight on (default 200 000)
void append_uint(std::string &s, uint64_t v);    // decimal digits, no allocation

template <typename F>
void for_each_member(const swbh_result &r, uint32_t sw, F &&f) {   // list order incl. grafted light swarms
  for (uint64_t k = 0; k < r.own_size[sw]; ++k) f(r.members[r.first[sw] + k]);
  for (uint32_t ls : r.grafted[sw])
    for (uint64_t k = 0; k < r.own_size[ls]; ++k) f(r.members[r.first[ls] + k]);
}
}  // namespace swb
