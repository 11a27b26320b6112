// swarm_b200/host/uclust.cc — the UCLUST-like writer (`-u`), the one output of swarm that needs alignments.
//
// Mirrors /root/reference src/algod1.cc:851-934 (d=1) and src/algo.cc:608-661 (d>1): per swarm one `C` and one `S`
// record, then one `H` record per other member carrying the global alignment of the member against the SEED
// (percent identity with one decimal, CIGAR of the alignment, "=" when there is no difference).
//
// The alignment is the reference's scalar Needleman-Wunsch-Sellers (src/nw.cc:40-191: affine gaps, costs
// mismatch / gap open / gap extension from src/swarm.cc:466-483, match = 0) including its tie-breaks, because the
// CIGAR depends on them: while tracing back from the lower-right corner an open gap is continued first, then a gap
// along the database sequence ("I"), then along the query ("D"), then the diagonal.  It is restated here over
// unpacked nucleotide bytes with the four decisions of a cell packed in one byte that is written once (the
// reference ORs flags into a zeroed matrix and clears it again after every pair), inside a band that is widened until
// the result provably equals the full matrix's (Aligner::align).  Pairs are independent, so the
// swarms are cut into contiguous ranges of equal alignment work and aligned by `threads` workers (the reference's
// `-t` — it aligns serially); the ranges' texts are concatenated in order, so the output does not depend on it.
#include "../../include/swarm_b200_host.h"
#include "amplicon_db.h"
#include "result.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace swb { int give_text(const std::string &s, char **out, uint64_t *out_len); }

namespace {

std::atomic<uint32_t> g_first_band{8};      // swbh_uclust_band(): first band half-width of the -u aligner (0 = full matrix)

constexpr uint8_t kUp = 1, kLeft = 2, kExtUp = 4, kExtLeft = 8;

struct Aligner {
  uint64_t mismatch, gap_open, gap_ext;
  std::vector<uint8_t> cell;          // rows (member) x columns (seed)
  std::vector<uint64_t> h, e;         // per column: best score ending here, best score ending in a gap along the member
  std::vector<uint8_t> d, q;          // unpacked nucleotides
  std::string ops;                    // alignment, traced from the end

  static void unpack(const uint64_t *w, uint32_t len, std::vector<uint8_t> &out) {
    out.resize(len);
    for (uint32_t i = 0; i < len; ++i) out[i] = static_cast<uint8_t>((w[i >> 5] >> ((i & 31u) * 2)) & 3u);
  }

  // One pass of the recurrence over the cells |column - row| <= w (everything outside counts as infinite), flags in
  // `cell[row][column - row + w]`.  Returns the corner score, or kInf when the corner is outside the band.
  static constexpr uint64_t kInf = 1ull << 60;
  uint64_t fill(uint32_t dlen, uint32_t qlen, uint32_t w) {
    const uint64_t go = gap_open, ge = gap_ext;
    const uint32_t nb = 2 * w + 1;
    cell.resize(static_cast<size_t>(dlen) * nb);
    h.resize(qlen); e.resize(qlen);
    for (uint32_t c = 0; c < qlen; ++c) {                        // the boundary row (src/nw.cc:71-75), inside its band
      h[c] = c <= w ? go + (c + 1) * ge : kInf;
      e[c] = c <= w ? 2 * go + (c + 2) * ge : kInf;
    }
    for (uint32_t r = 0; r < dlen; ++r) {
      const uint32_t c_lo = r > w ? r - w : 0, c_hi = std::min<uint64_t>(qlen - 1, static_cast<uint64_t>(r) + w);
      if (c_lo > c_hi) return kInf;
      uint64_t up, diag;
      if (r <= w) { up = 2 * go + (r + 2) * ge; diag = r == 0 ? 0 : go + r * ge; }   // true left boundary (src/nw.cc:75-76)
      else { up = kInf; diag = h[c_lo - 1]; }                    // cell (r, c_lo - 1) is outside the band
      uint8_t *row = cell.data() + static_cast<size_t>(r) * nb + w - r;      // row[c] = cell[r][c - r + w]
      const uint8_t dn = d[r];
      for (uint32_t c = c_lo; c <= c_hi; ++c) {
        const uint64_t h_above = h[c];
        uint64_t left = e[c];
        uint64_t best = diag + (dn == q[c] ? 0 : mismatch);
        // the four decisions of the cell as flag arithmetic (no data-dependent branches)
        uint32_t f = static_cast<uint32_t>(up < best);                                  // kUp
        best = std::min(best, std::min(up, left));
        f |= static_cast<uint32_t>(left == best) << 1;                                  // kLeft
        h[c] = best;
        const uint64_t opened = best + go + ge;
        left += ge; up += ge;
        f |= static_cast<uint32_t>(up < opened) << 2;                                   // kExtUp
        f |= static_cast<uint32_t>(left < opened) << 3;                                 // kExtLeft
        up = std::min(up, opened);
        e[c] = std::min(left, opened);
        row[c] = static_cast<uint8_t>(f);
        diag = h_above;
      }
    }
    const uint32_t gap = qlen > dlen ? qlen - dlen : dlen - qlen;
    return gap <= w ? h[qlen - 1] : kInf;
  }

  // fills `ops` (reversed) and returns the number of alignment columns that are not matches.
  // The matrix is filled inside a band that is widened until the result is provably the full matrix's: k gap columns
  // cost at least gap_open + k * gap_ext, so an alignment that costs less than gap_open + (w + 1) * gap_ext never
  // leaves the diagonals +-w; if the banded corner score is below that, so is the optimum, every optimal path — and
  // every predecessor that ties with or beats a step of one, which is what the flags record — lies inside the band,
  // and the trace-back reads the flags the full matrix would hold (same argument as the device aligner, DESIGN.md
  // §3.6).  Members of a swarm are a few edits from their seed: 150 x 17 cells instead of 150 x 150.
  uint64_t align(const uint64_t *member, uint32_t dlen, const uint64_t *seed, uint32_t qlen) {
    unpack(member, dlen, d);
    unpack(seed, qlen, q);
    const uint32_t full = std::max(dlen, qlen);
    uint32_t w = first_band == 0 ? full : std::min(first_band, full);
    for (;;) {
      const uint64_t score = fill(dlen, qlen, w);
      if (w >= full || score < gap_open + (static_cast<uint64_t>(w) + 1) * gap_ext) break;
      // the banded score bounds the optimum from above: the band that makes THAT score provable ends the search in one
      // more pass (a wider band can only lower the score); no score at all (corner outside the band): four times wider
      const uint64_t need = score < kInf ? (score - gap_open) / gap_ext + 1 : 4ull * w;
      w = static_cast<uint32_t>(std::min<uint64_t>(full, std::max<uint64_t>(need, static_cast<uint64_t>(w) + 1)));
    }
    const uint32_t nb = 2 * w + 1;
    // trace back (src/nw.cc:111-191)
    ops.clear();
    uint64_t matches = 0;
    uint32_t c = qlen, r = dlen;
    char op = 0;
    while (c > 0 && r > 0) {
      const uint8_t f = cell[static_cast<size_t>(r - 1) * nb + (c - 1) + w - (r - 1)];
      if (op == 'I' && (f & kExtLeft)) { --r; }
      else if (op == 'D' && (f & kExtUp)) { --c; }
      else if (f & kLeft) { --r; op = 'I'; }
      else if (f & kUp) { --c; op = 'D'; }
      else { --r; --c; op = 'M'; if (q[c] == d[r]) ++matches; }
      ops.push_back(op);
    }
    ops.append(c, 'D');
    ops.append(r, 'I');
    return ops.size() - matches;
  }
  uint32_t first_band = 8;
};

// run-length form of the alignment read from its far end (src/utils/cigar.cc:30-60: a count of 1 is not printed)
void append_cigar_reversed(std::string &out, const std::string &ops) {
  for (size_t i = ops.size(); i > 0;) {
    const char op = ops[i - 1];
    size_t run = 0;
    while (i > 0 && ops[i - 1] == op) { --i; ++run; }
    if (run > 1) out += std::to_string(run);
    out += op;
  }
}

struct Unit { uint32_t swarm, cluster_no; };

}  // namespace

extern "C" void swbh_uclust_band(int first_half_width) { g_first_band.store(first_half_width < 0 ? 0u : static_cast<uint32_t>(first_half_width)); }

extern "C" int swbh_write_uclust(const swbh_db *dbh, const swbh_result *r, int64_t differences, const int64_t penalties[3],
                                 int usearch, int64_t append, int threads, char **out, uint64_t *out_len) {
  const swb::AmpliconDb &db = dbh->db;
  swb::DbOptions o; o.usearch_abundance = usearch != 0; o.append_abundance = append;
  const bool dn = differences > 1;
  // d>1 lists the hits in the order the greedy loop found them (src/algo.cc:621: hits[1..]): (sub)seeds in list
  // order, each one's hits by ascending id — the order of the `-i` lines; d=1 walks the swarm's list (:888)
  std::vector<uint64_t> cstart;
  std::vector<uint32_t> child;
  if (dn) {
    const uint32_t n = r->n;
    cstart.assign(static_cast<size_t>(n) + 1, 0);
    for (uint32_t a = 0; a < n; ++a) if (r->parent[a] != 0xFFFFFFFFu) cstart[r->parent[a] + 1]++;
    for (uint32_t a = 0; a < n; ++a) cstart[a + 1] += cstart[a];
    child.resize(cstart[n]);
    std::vector<uint64_t> cur(cstart.begin(), cstart.end() - 1);
    for (uint32_t a = 0; a < n; ++a) if (r->parent[a] != 0xFFFFFFFFu) child[cur[r->parent[a]]++] = a;
  }
  std::vector<Unit> units;
  uint64_t work = 0;
  for (uint32_t sw = 0; sw < r->seed.size(); ++sw) {
    if (r->attached[sw]) continue;
    units.push_back({sw, static_cast<uint32_t>(units.size())});
    work += r->size[sw];
  }
  const unsigned T = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>({threads, 512, static_cast<int64_t>(units.size())})));
  std::vector<std::string> text(T);
  std::vector<size_t> cut(T + 1, units.size());
  cut[0] = 0;
  {
    uint64_t acc = 0;
    unsigned t = 1;
    for (size_t u = 0; u < units.size() && t < T; ++u) {
      acc += r->size[units[u].swarm];
      if (acc * T >= work * t) cut[t++] = u + 1;
    }
  }
  auto run = [&](unsigned t) {
    Aligner A;
    A.first_band = g_first_band.load();
    A.mismatch = static_cast<uint64_t>(penalties[0]); A.gap_open = static_cast<uint64_t>(penalties[1]); A.gap_ext = static_cast<uint64_t>(penalties[2]);
    std::string &s = text[t];
    char num[64];
    for (size_t u = cut[t]; u < cut[t + 1]; ++u) {
      const uint32_t sw = units[u].swarm, no = units[u].cluster_no, seed = r->seed[sw];
      s += "C\t" + std::to_string(no) + "\t" + std::to_string(r->size[sw]) + "\t*\t*\t*\t*\t*\t";
      swb::append_id(s, db, seed, o);
      s += "\t*\nS\t" + std::to_string(no) + "\t" + std::to_string(db.len[seed]) + "\t*\t*\t*\t*\t*\t";
      swb::append_id(s, db, seed, o);
      s += "\t*\n";
      auto hit = [&](uint32_t a) {
        if (a == seed) return;
        const uint64_t diff = A.align(db.seq(a), db.len[a], db.seq(seed), db.len[seed]);
        const double alen = static_cast<double>(A.ops.size());
        std::snprintf(num, sizeof num, "%.1f", 100.0 * (alen - static_cast<double>(diff)) / alen);
        s += "H\t" + std::to_string(no) + "\t" + std::to_string(db.len[a]) + "\t" + num + "\t+\t0\t0\t";
        if (diff > 0) append_cigar_reversed(s, A.ops); else s += '=';
        s += '\t';
        swb::append_id(s, db, a, o);
        s += '\t';
        swb::append_id(s, db, seed, o);
        s += '\n';
      };
      if (dn) {
        for (uint64_t k = 0; k < r->own_size[sw]; ++k) {
          const uint32_t par = r->members[r->first[sw] + k];
          for (uint64_t c = cstart[par]; c < cstart[par + 1]; ++c) hit(child[c]);
        }
      } else {
        swb::for_each_member(*r, sw, hit);
      }
    }
  };
  if (T == 1) run(0);
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(run, t);
    for (auto &th : pool) th.join();
  }
  std::string all;
  size_t total = 0;
  for (const auto &s : text) total += s.size();
  all.reserve(total);
  for (const auto &s : text) all += s;
  return swb::give_text(all, out, out_len);
}
