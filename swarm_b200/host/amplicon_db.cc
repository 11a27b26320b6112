// swarm_b200/host/amplicon_db.cc — FASTA → sorted, 2-bit packed, fixed-stride SoA database.
// Behavioural mirror of /root/reference src/db.cc (citations inline); written from scratch.
//
// Two parsers produce the same database.  db_parse_serial() is the statement of the rules: it walks the text once,
// in order, and owns every error message.  db_parse_parallel() is the ingest path for large inputs (SURVEY.md §8 row
// f1: at 10 M amplicons the reference spends ~12 s here on one core, three orders of magnitude more than the GPU
// spends clustering): the text is cut at record starts into one range per worker, every worker packs its records,
// abundance annotations are parsed per record, duplicate labels (and, for d > 1, duplicate sequences) are found with
// a lock-free open-addressing set of record indices, the order is a parallel merge sort and the SoA gather is split
// by rows.  It only ever COMPLETES on well-formed input: anything irregular (a NUL byte, an illegal character, a
// missing annotation, a duplicate ...) makes it give up, and the serial parser then reports the reference's message.
#include "amplicon_db.h"

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <numeric>
#include <thread>
#include <unordered_set>

namespace swb {
namespace {

constexpr uint32_t kMaxSequenceLength = 67108861;   // src/db.cc:439
constexpr uint32_t kMaxHeaderLength = 16777215;     // src/db.cc:443

struct RawEntry {
  uint64_t header_pos;   // offset of the header text in the input
  uint32_t header_len;
  uint64_t word_pos;     // first packed word in `packed`
  uint32_t len;
  uint32_t lineno;
  uint64_t abundance;
  int32_t ab_start, ab_end;
};

// (_)([0-9]+)$ on the LAST underscore — src/db.cc:161-211
bool find_swarm_abundance(const char *h, uint32_t hlen, int32_t &start, int32_t &end, int64_t &number) {
  start = end = 0; number = 0;
  int64_t us = -1;
  for (int64_t i = static_cast<int64_t>(hlen) - 1; i >= 0; --i) if (h[i] == '_') { us = i; break; }
  if (us < 0) return false;
  uint32_t nd = 0;
  while (us + 1 + nd < hlen && h[us + 1 + nd] >= '0' && h[us + 1 + nd] <= '9') ++nd;
  if (nd > 20) return false;
  if (us + 1 + nd != hlen) return false;
  start = static_cast<int32_t>(us);
  end = static_cast<int32_t>(us + 1 + nd);
  number = nd ? std::atol(std::string(h + us + 1, nd).c_str()) : 0;   // atol semantics as the reference
  return true;
}

// (^|;)size=([0-9]+)(;|$) — src/db.cc:214-283
bool find_usearch_abundance(const char *h, uint32_t hlen_u, int32_t &start, int32_t &end, int64_t &number) {
  start = end = 0; number = 0;
  const int64_t hlen = hlen_u, alen = 5;
  const std::string hs(h, hlen_u);
  int64_t position = 0;
  while (position + alen < hlen) {
    const size_t r = hs.find("size=", static_cast<size_t>(position));
    if (r == std::string::npos) break;
    position = static_cast<int64_t>(r);
    if (position > 0 && h[position - 1] != ';') { position += alen + 1; continue; }
    int64_t nd = 0;
    while (position + alen + nd < hlen && h[position + alen + nd] >= '0' && h[position + alen + nd] <= '9') ++nd;
    if (nd == 0) { position += alen + 1; continue; }
    if (position + alen + nd < hlen && h[position + alen + nd] != ';') { position += alen + nd + 2; continue; }
    start = position > 0 ? static_cast<int32_t>(position - 1) : 0;
    end = static_cast<int32_t>(std::min(position + alen + nd + 1, hlen));
    number = std::atol(hs.substr(static_cast<size_t>(position + alen), static_cast<size_t>(nd)).c_str());
    return true;
  }
  return false;
}

inline int map_nt(unsigned char c) {   // src/db.cc:100-114: 1-based code, 0 = not a nucleotide
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': case 'U': case 'u': return 4;
    default: return 0;
  }
}

}  // namespace

std::string db_parse_serial(const char *text, uint64_t size, const DbOptions &opt, AmpliconDb &db) {
  db = AmpliconDb{};
  std::vector<RawEntry> entries;
  std::vector<uint64_t> packed;
  packed.reserve(size / 28 + 16);
  uint64_t pos = 0;
  uint32_t lineno = 1;
  uint64_t missing = 0; uint32_t missing_lineno = 0; uint64_t missing_entry = 0;
  char msg[512];

  auto line_end = [&](uint64_t p) {   // one past the '\n' (or size)
    const void *nl = std::memchr(text + p, '\n', size - p);
    return nl ? static_cast<uint64_t>(static_cast<const char *>(nl) - text) + 1 : size;
  };

  while (pos < size && text[pos] != '\0') {
    if (text[pos] != '>') return "Illegal header line in fasta file.";            // src/db.cc:492-494
    uint64_t le = line_end(pos);
    RawEntry e{};
    e.header_pos = pos + 1;
    uint32_t hl = 0;
    while (pos + 1 + hl < le) {                                                   // src/db.cc:498-499
      const char c = text[pos + 1 + hl];
      if (c == ' ' || c == '\r' || c == '\n' || c == '\0') break;
      ++hl;
    }
    e.header_len = hl;
    db.longest_header = std::max(db.longest_header, hl);
    if (db.longest_header > kMaxHeaderLength) return "Headers longer than 16,777,215 symbols are not supported.";
    e.lineno = lineno;
    pos = le;
    ++lineno;
    e.word_pos = packed.size();
    uint64_t buf = 0; uint32_t nbuf = 0, length = 0;
    while (pos < size && text[pos] != '\0' && text[pos] != '>') {                 // src/db.cc:555-603
      le = line_end(pos);
      for (uint64_t p = pos; p < le; ++p) {
        const unsigned char c = static_cast<unsigned char>(text[p]);
        if (c == '\0') break;                                                     // getline C-string semantics
        const int m = map_nt(c);
        if (m != 0) {
          buf |= static_cast<uint64_t>(m - 1) << (2 * nbuf);
          ++length;
          if (++nbuf == 32) { packed.push_back(buf); buf = 0; nbuf = 0; }
        } else if (c != '\n' && c != '\r') {
          if (c >= 32 && c <= 126)
            std::snprintf(msg, sizeof msg, "Illegal character '%c' in sequence on line %u.", c, lineno);
          else
            std::snprintf(msg, sizeof msg, "Illegal character (ascii no %u) in sequence on line %u.", c, lineno);
          return msg;
        }
      }
      if (length > kMaxSequenceLength) return "Sequences longer than 67,108,861 symbols are not supported.";
      pos = le;
      ++lineno;
    }
    if (length == 0) {                                                            // src/db.cc:608-611
      std::snprintf(msg, sizeof msg, "Empty sequence found on line %u.", lineno - 1);
      return msg;
    }
    if (nbuf > 0) packed.push_back(buf);
    e.len = length;
    db.nucleotides += length;
    db.longest = std::max(db.longest, length);
    entries.push_back(e);
  }

  const uint64_t n64 = entries.size();
  if (n64 >= 0xFFFFFFFFull) return "Too many sequences (amplicon ids are 32-bit).";
  const uint32_t n = static_cast<uint32_t>(n64);

  // abundance annotations, empty identifiers, duplicated identifiers — src/db.cc:286-347,676-758
  std::unordered_set<std::string> labels;
  labels.reserve(static_cast<size_t>(n) * 2);
  for (uint32_t i = 0; i < n; ++i) {
    RawEntry &e = entries[i];
    const char *h = text + e.header_pos;
    int64_t number = 0; int32_t s = 0, t = 0;
    int64_t abundance = 0;
    const bool found = opt.usearch_abundance ? find_usearch_abundance(h, e.header_len, s, t, number)
                                             : find_swarm_abundance(h, e.header_len, s, t, number);
    if (found) {
      if (number <= 0) {
        return "Illegal abundance value on line " + std::to_string(e.lineno) + ":\n" + std::string(h, e.header_len) +
               "\nAbundance values should be positive integers.";
      }
      abundance = number;
    }
    if (abundance == 0) {
      s = t = static_cast<int32_t>(e.header_len);
      if (opt.append_abundance != 0) abundance = opt.append_abundance;
      else if (++missing == 1) { missing_lineno = e.lineno; missing_entry = i; }
    }
    e.abundance = static_cast<uint64_t>(abundance);
    e.ab_start = s; e.ab_end = t;
    if (s == 0 && t == static_cast<int32_t>(e.header_len)) return "Empty sequence identifier.";
    const int32_t id_start = s > 0 ? 0 : t;
    const int32_t id_len = s > 0 ? s : static_cast<int32_t>(e.header_len) - t;
    std::string label(h + id_start, static_cast<size_t>(id_len));
    if (!labels.insert(label).second) return "Duplicated sequence identifier: " + label;
  }
  labels.clear();

  if (opt.check_duplicate_sequences) {                                            // src/db.cc:763-796
    std::unordered_set<std::string> seqs;
    seqs.reserve(static_cast<size_t>(n) * 2);
    for (uint32_t i = 0; i < n; ++i) {
      const RawEntry &e = entries[i];
      std::string key(reinterpret_cast<const char *>(packed.data() + e.word_pos), ((e.len + 31) / 32) * 8);
      key.append(reinterpret_cast<const char *>(&e.len), 4);
      if (!seqs.insert(std::move(key)).second)
        return "some fasta entries have identical sequences.\n"
               "Swarm expects dereplicated fasta files.\n"
               "Such files can be produced with swarm or vsearch:\n"
               " swarm -d 0 -w derep.fasta -o /dev/null input.fasta\n"
               "or\n"
               " vsearch --derep_fulllength input.fasta --sizein --sizeout --output derep.fasta";
    }
  }

  if (missing != 0) {                                                             // src/db.cc:374-385
    const RawEntry &e = entries[missing_entry];
    return "Abundance annotations not found for " + std::to_string(missing) + " sequences, starting on line " +
           std::to_string(missing_lineno) + ".\n>" + std::string(text + e.header_pos, e.header_len) + "\n" +
           "Fasta headers must end with abundance annotations (_INT or ;size=INT).\n"
           "The -z option must be used if the abundance annotation is in the latter format.\n"
           "Abundance annotations can be produced by dereplicating the sequences.\n"
           "The header is defined as the string comprised between the \">\" symbol\n"
           "and the first space or the end of the line, whichever comes first.";
  }

  // sort: abundance descending, then strcmp(full header) ascending — src/db.cc:392-411
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  auto less = [&](uint32_t a, uint32_t b) {
    const RawEntry &x = entries[a], &y = entries[b];
    if (x.abundance != y.abundance) return x.abundance > y.abundance;
    const uint32_t m = std::min(x.header_len, y.header_len);
    const int c = std::memcmp(text + x.header_pos, text + y.header_pos, m);   // headers hold no NUL: == strcmp
    if (c != 0) return c < 0;
    return x.header_len < y.header_len;
  };
  if (!std::is_sorted(order.begin(), order.end(), less)) std::sort(order.begin(), order.end(), less);

  // gather into the SoA layout
  db.n = n;
  db.stride = std::max<uint32_t>(1, (db.longest + 31) / 32);
  db.words.assign(static_cast<uint64_t>(n) * db.stride, 0);
  db.len.resize(n); db.abundance.resize(n); db.ab_start.resize(n); db.ab_end.resize(n);
  db.header_off.resize(static_cast<uint64_t>(n) + 1);
  uint64_t hbytes = 0;
  for (uint32_t i = 0; i < n; ++i) hbytes += entries[i].header_len + 1;
  db.headers.resize(hbytes);
  uint64_t hp = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const RawEntry &e = entries[order[i]];
    db.len[i] = e.len; db.abundance[i] = e.abundance; db.ab_start[i] = e.ab_start; db.ab_end[i] = e.ab_end;
    std::memcpy(db.words.data() + static_cast<uint64_t>(i) * db.stride, packed.data() + e.word_pos, ((e.len + 31) / 32) * 8);
    db.header_off[i] = hp;
    std::memcpy(db.headers.data() + hp, text + e.header_pos, e.header_len);
    db.headers[hp + e.header_len] = '\0';
    hp += e.header_len + 1;
  }
  db.header_off[n] = hp;
  return "";
}

// ------------------------------------------------------------------------------------------------------------------
// parallel ingest
// ------------------------------------------------------------------------------------------------------------------
namespace {

std::atomic<int> g_threads{0};           // 0 = hardware concurrency (at most 32)

unsigned ingest_threads() {
  int t = g_threads.load();
  if (t <= 0) t = static_cast<int>(std::min(32u, std::max(1u, std::thread::hardware_concurrency())));
  return static_cast<unsigned>(std::min(t, 512));       // the sample sort keeps bucket ids in 16 bits and samples 64 x T records
}

template <typename F>
void run_workers(unsigned T, F &&f) {
  if (T <= 1) { f(0u); return; }
  std::vector<std::thread> pool;
  pool.reserve(T);
  for (unsigned t = 0; t < T; ++t) pool.emplace_back([&f, t] { f(t); });
  for (auto &th : pool) th.join();
}

struct ParEntry {
  uint64_t header_pos;
  uint64_t word_pos;         // first packed word inside the buffer of worker `range`
  uint64_t abundance;
  uint32_t header_len, len;
  int32_t ab_start, ab_end;
  uint32_t range;
};

struct Range {
  uint64_t begin = 0, end = 0;
  std::vector<ParEntry> entries;
  std::vector<uint64_t> packed;
  uint64_t nucleotides = 0;
  uint32_t longest = 0, longest_header = 0;
  bool ok = true;
};

// 0..3 nucleotide code, 4 = line break (skipped), 5 = anything else (src/db.cc:100-114,555-603)
struct NtTable {
  uint8_t v[256];
  NtTable() {
    for (auto &x : v) x = 5;
    v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = v['U'] = v['u'] = 3;
    v['\n'] = v['\r'] = 4;
  }
};
const NtTable kNt;

void parse_range(const char *text, Range &R, uint32_t range_no) {
  uint64_t pos = R.begin;
  const uint64_t end = R.end;
  R.packed.reserve((end - pos) / 28 + 16);
  R.entries.reserve((end - pos) / 160 + 16);
  while (pos < end) {
    if (text[pos] != '>') { R.ok = false; return; }
    const void *nl = std::memchr(text + pos, '\n', end - pos);
    uint64_t le = nl ? static_cast<uint64_t>(static_cast<const char *>(nl) - text) + 1 : end;
    if (std::memchr(text + pos, '\0', le - pos) != nullptr) { R.ok = false; return; }   // NUL bytes: the serial parser's business
    ParEntry e{};
    e.range = range_no;
    e.header_pos = pos + 1;
    uint32_t hl = 0;
    while (pos + 1 + hl < le) {
      const char c = text[pos + 1 + hl];
      if (c == ' ' || c == '\r' || c == '\n') break;
      ++hl;
    }
    if (hl > kMaxHeaderLength) { R.ok = false; return; }
    e.header_len = hl;
    R.longest_header = std::max(R.longest_header, hl);
    pos = le;
    const size_t first_word = R.packed.size();
    uint64_t buf = 0, length = 0;
    uint32_t nbuf = 0;
    // the sequence runs to the next line that starts with '>' (a '>' elsewhere is an illegal character).  Line by line:
    // 32 nucleotides at a time go into one word with no per-character branch (the codes of a block are OR-ed and looked
    // at once); a block holding anything but nucleotides — a line break, '\r', an illegal byte — takes the slow lane.
    while (pos < end && text[pos] != '>') {
      const void *lf = std::memchr(text + pos, '\n', end - pos);
      const uint64_t stop = lf ? static_cast<uint64_t>(static_cast<const char *>(lf) - text) + 1 : end;
      while (pos < stop) {
        if (nbuf == 0 && stop - pos >= 32) {
          uint64_t word = 0;
          uint32_t seen = 0;
          for (uint32_t k = 0; k < 32; ++k) {
            const uint32_t code = kNt.v[static_cast<unsigned char>(text[pos + k])];
            seen |= code;
            word |= static_cast<uint64_t>(code & 3u) << (2 * k);
          }
          if (seen < 4) { R.packed.push_back(word); length += 32; pos += 32; continue; }
        }
        const uint8_t code = kNt.v[static_cast<unsigned char>(text[pos])];
        if (code < 4) {
          buf |= static_cast<uint64_t>(code) << (2 * nbuf);
          ++length;
          if (++nbuf == 32) { R.packed.push_back(buf); buf = 0; nbuf = 0; }
        } else if (code == 5) { R.ok = false; return; }
        ++pos;
      }
    }
    if (length == 0 || length > kMaxSequenceLength) { R.ok = false; return; }
    if (nbuf > 0) R.packed.push_back(buf);
    e.len = static_cast<uint32_t>(length);
    e.word_pos = first_word;
    R.nucleotides += length;
    R.longest = std::max(R.longest, e.len);
    R.entries.push_back(e);
  }
}

inline uint64_t hash_bytes(const void *p, size_t n, uint64_t seed) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  uint64_t h = seed ^ (n * 0x9e3779b97f4a7c15ull);
  while (n >= 8) {
    uint64_t w;
    std::memcpy(&w, b, 8);
    h = (h ^ w) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    b += 8; n -= 8;
  }
  uint64_t w = 0;
  std::memcpy(&w, b, n);
  h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
  return h ^ (h >> 29);
}

// a key of the duplicate checks: bytes + a tag that must also match (the sequence length: "A" and "AA" pack to equal words)
struct Key { const void *p; size_t n; uint32_t tag; };

// lock-free open-addressing set of record indices: false as soon as two records carry equal keys
template <typename KeyOf>
bool all_distinct(uint32_t n, unsigned T, KeyOf &&key_of) {
  uint64_t slots = 16;
  while (slots < static_cast<uint64_t>(n) * 2) slots <<= 1;
  std::vector<std::atomic<uint32_t>> table(slots);
  std::vector<uint64_t> hashes(n);
  std::atomic<bool> dup{false};
  run_workers(T, [&](unsigned t) {
    const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = static_cast<uint64_t>(n) * (t + 1) / T;
    for (uint64_t s = slots * t / T; s < slots * (t + 1) / T; ++s) table[s].store(0xFFFFFFFFu, std::memory_order_relaxed);
    for (uint64_t i = lo; i < hi; ++i) { const Key k = key_of(static_cast<uint32_t>(i)); hashes[i] = hash_bytes(k.p, k.n, k.tag); }
  });
  run_workers(T, [&](unsigned t) {
    const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = static_cast<uint64_t>(n) * (t + 1) / T;
    for (uint64_t i = lo; i < hi && !dup.load(std::memory_order_relaxed); ++i) {
      const Key k = key_of(static_cast<uint32_t>(i));
      uint64_t j = hashes[i] & (slots - 1);
      for (;;) {
        uint32_t cur = table[j].load(std::memory_order_acquire);
        if (cur == 0xFFFFFFFFu && table[j].compare_exchange_strong(cur, static_cast<uint32_t>(i), std::memory_order_acq_rel)) break;
        if (hashes[cur] == hashes[i]) {
          const Key o = key_of(cur);
          if (o.tag == k.tag && o.n == k.n && std::memcmp(o.p, k.p, k.n) == 0) { dup.store(true); break; }
        }
        j = (j + 1) & (slots - 1);
      }
    }
  });
  return !dup.load();
}

// sample sort: T-1 splitters from a sorted sample, one counting + one scattering pass over the workers' slices, then
// every bucket is sorted by its own worker — no merge passes (the keys are unique, so the buckets are balanced)
template <typename Rec, typename Less>
void parallel_sort(raw_vector<Rec> &v, unsigned T, Less less) {
  const size_t n = v.size();
  if (T <= 1 || n < (1u << 16)) { std::sort(v.begin(), v.end(), less); return; }
  const size_t per_bucket = 64, S = per_bucket * T;
  std::vector<Rec> sample(S);
  for (size_t i = 0; i < S; ++i) sample[i] = v[i * (n / S)];
  std::sort(sample.begin(), sample.end(), less);
  std::vector<Rec> split(T - 1);
  for (unsigned b = 1; b < T; ++b) split[b - 1] = sample[b * per_bucket];
  raw_vector<uint16_t> bucket(n);
  std::vector<size_t> count(static_cast<size_t>(T) * T, 0);        // [worker][bucket]
  run_workers(T, [&](unsigned t) {
    const size_t lo = n * t / T, hi = n * (t + 1) / T;
    size_t *c = count.data() + static_cast<size_t>(t) * T;
    for (size_t i = lo; i < hi; ++i) {
      const unsigned b = static_cast<unsigned>(std::upper_bound(split.begin(), split.end(), v[i], less) - split.begin());
      bucket[i] = static_cast<uint16_t>(b);
      ++c[b];
    }
  });
  std::vector<size_t> start(static_cast<size_t>(T) * T), bucket_begin(T + 1, 0);
  size_t run = 0;
  for (unsigned b = 0; b < T; ++b) {
    bucket_begin[b] = run;
    for (unsigned t = 0; t < T; ++t) { start[static_cast<size_t>(t) * T + b] = run; run += count[static_cast<size_t>(t) * T + b]; }
  }
  bucket_begin[T] = run;
  raw_vector<Rec> out(n);
  run_workers(T, [&](unsigned t) {
    const size_t lo = n * t / T, hi = n * (t + 1) / T;
    size_t *at = start.data() + static_cast<size_t>(t) * T;
    for (size_t i = lo; i < hi; ++i) out[at[bucket[i]]++] = v[i];
  });
  run_workers(T, [&](unsigned b) {
    std::sort(out.begin() + static_cast<int64_t>(bucket_begin[b]), out.begin() + static_cast<int64_t>(bucket_begin[b + 1]), less);
  });
  v.swap(out);
}

// returns false when the input is not plainly well-formed (the serial parser then decides and words the error)
// SWARM_B200_INGEST_TIMES=1: stage times of the parallel ingest on stderr (diagnostic)
struct StageClock {
  bool on = std::getenv("SWARM_B200_INGEST_TIMES") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ingest] %-9s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};

bool db_parse_parallel(const char *text, uint64_t size, const DbOptions &opt, AmpliconDb &db, unsigned T) {
  StageClock clock;
  if (size == 0) return false;
  // ranges start at a '>' that follows a line break
  std::vector<Range> ranges(T);
  {
    uint64_t prev = 0;
    for (unsigned t = 0; t < T; ++t) {
      uint64_t want = t + 1 == T ? size : size * (t + 1) / T;
      if (want < prev) want = prev;
      while (want < size && !(want > 0 && text[want] == '>' && text[want - 1] == '\n')) {
        const void *g = std::memchr(text + want + 1, '>', size - want - 1);
        want = g ? static_cast<uint64_t>(static_cast<const char *>(g) - text) : size;
      }
      ranges[t].begin = prev; ranges[t].end = want;
      prev = want;
    }
  }
  run_workers(T, [&](unsigned t) { parse_range(text, ranges[t], t); });
  clock.lap("pack");
  uint64_t n64 = 0;
  for (const Range &R : ranges) { if (!R.ok) return false; n64 += R.entries.size(); }
  if (n64 == 0 || n64 >= 0xFFFFFFF0ull) return false;
  const uint32_t n = static_cast<uint32_t>(n64);
  raw_vector<ParEntry> entries(n);
  db = AmpliconDb{};
  std::vector<uint64_t> first(T + 1, 0);
  for (unsigned t = 0; t < T; ++t) {
    const Range &R = ranges[t];
    first[t + 1] = first[t] + R.entries.size();
    db.nucleotides += R.nucleotides;
    db.longest = std::max(db.longest, R.longest);
    db.longest_header = std::max(db.longest_header, R.longest_header);
  }
  run_workers(T, [&](unsigned t) {
    std::copy(ranges[t].entries.begin(), ranges[t].entries.end(), entries.begin() + static_cast<int64_t>(first[t]));
    std::vector<ParEntry>().swap(ranges[t].entries);
  });
  clock.lap("merge");
  // abundance annotations and identifiers (src/db.cc:286-347,676-758): any irregular record ends the fast path
  std::atomic<bool> bad{false};
  run_workers(T, [&](unsigned t) {
    const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = static_cast<uint64_t>(n) * (t + 1) / T;
    for (uint64_t i = lo; i < hi; ++i) {
      ParEntry &e = entries[i];
      const char *h = text + e.header_pos;
      int64_t number = 0; int32_t s = 0, u = 0;
      const bool found = opt.usearch_abundance ? find_usearch_abundance(h, e.header_len, s, u, number)
                                               : find_swarm_abundance(h, e.header_len, s, u, number);
      int64_t abundance = 0;
      if (found) { if (number <= 0) { bad.store(true); return; } abundance = number; }
      if (abundance == 0) {
        s = u = static_cast<int32_t>(e.header_len);
        if (opt.append_abundance == 0) { bad.store(true); return; }
        abundance = opt.append_abundance;
      }
      if (s == 0 && u == static_cast<int32_t>(e.header_len)) { bad.store(true); return; }
      e.abundance = static_cast<uint64_t>(abundance);
      e.ab_start = s; e.ab_end = u;
    }
  });
  if (bad.load()) return false;
  clock.lap("abundance");
  auto words_of = [&](const ParEntry &e) { return ranges[e.range].packed.data() + e.word_pos; };
  auto label_of = [&](uint32_t i) {
    const ParEntry &e = entries[i];
    const int32_t id_start = e.ab_start > 0 ? 0 : e.ab_end;
    const int32_t id_len = e.ab_start > 0 ? e.ab_start : static_cast<int32_t>(e.header_len) - e.ab_end;
    return Key{text + e.header_pos + id_start, static_cast<size_t>(id_len), 0u};
  };
  if (!all_distinct(n, T, label_of)) return false;
  if (opt.check_duplicate_sequences) {                                            // src/db.cc:763-796 (d > 1)
    auto seq_of = [&](uint32_t i) {
      const ParEntry &e = entries[i];
      return Key{words_of(e), static_cast<size_t>((e.len + 31) / 32) * 8, e.len};
    };
    if (!all_distinct(n, T, seq_of)) return false;
  }
  // order: abundance descending, then header ascending (src/db.cc:392-411)
  clock.lap("distinct");
  // sort records carry the abundance and the first 8 header bytes (big-endian, zero padded: a shorter header that is
  // a prefix sorts first, like strcmp) so that almost every comparison is decided without touching the text
  struct SortRec { uint64_t abundance, prefix; uint32_t idx; };
  raw_vector<SortRec> order(n);
  auto less = [&](const SortRec &a, const SortRec &b) {
    if (a.abundance != b.abundance) return a.abundance > b.abundance;
    if (a.prefix != b.prefix) return a.prefix < b.prefix;
    const ParEntry &x = entries[a.idx], &y = entries[b.idx];
    const uint32_t m = std::min(x.header_len, y.header_len);
    const int c = std::memcmp(text + x.header_pos, text + y.header_pos, m);
    if (c != 0) return c < 0;
    return x.header_len < y.header_len;
  };
  {
    run_workers(T, [&](unsigned t) {
      const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = static_cast<uint64_t>(n) * (t + 1) / T;
      for (uint64_t i = lo; i < hi; ++i) {
        const ParEntry &e = entries[i];
        uint64_t pfx = 0;
        for (uint32_t k = 0; k < 8; ++k) pfx = (pfx << 8) | (k < e.header_len ? static_cast<unsigned char>(text[e.header_pos + k]) : 0u);
        order[i] = SortRec{e.abundance, pfx, static_cast<uint32_t>(i)};
      }
    });
    std::atomic<bool> sorted{true};
    run_workers(T, [&](unsigned t) {
      const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = std::min<uint64_t>(n - 1, static_cast<uint64_t>(n) * (t + 1) / T);
      for (uint64_t i = lo; i < hi && sorted.load(std::memory_order_relaxed); ++i)
        if (less(order[i + 1], order[i])) sorted.store(false);
    });
    clock.lap("sort keys");
    if (!sorted.load()) parallel_sort(order, T, less);
  }
  clock.lap("sort");
  // gather into the SoA layout
  db.n = n;
  db.stride = std::max<uint32_t>(1, (db.longest + 31) / 32);
  db.words.resize(static_cast<uint64_t>(n) * db.stride);          // uninitialised: every row is written whole below
  db.len.resize(n); db.abundance.resize(n); db.ab_start.resize(n); db.ab_end.resize(n);
  db.header_off.resize(static_cast<uint64_t>(n) + 1);
  std::vector<uint64_t> part(T + 1, 0);
  run_workers(T, [&](unsigned t) {
    const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = static_cast<uint64_t>(n) * (t + 1) / T;
    uint64_t b = 0;
    for (uint64_t i = lo; i < hi; ++i) b += entries[order[i].idx].header_len + 1;
    part[t + 1] = b;
  });
  for (unsigned t = 0; t < T; ++t) part[t + 1] += part[t];
  db.headers.resize(part[T]);
  run_workers(T, [&](unsigned t) {
    const uint64_t lo = static_cast<uint64_t>(n) * t / T, hi = static_cast<uint64_t>(n) * (t + 1) / T;
    uint64_t hp = part[t];
    for (uint64_t i = lo; i < hi; ++i) {
      const ParEntry &e = entries[order[i].idx];
      db.len[i] = e.len; db.abundance[i] = e.abundance; db.ab_start[i] = e.ab_start; db.ab_end[i] = e.ab_end;
      const uint32_t nw = (e.len + 31) / 32;
      uint64_t *row = db.words.data() + i * db.stride;
      std::memcpy(row, words_of(e), static_cast<size_t>(nw) * 8);
      std::fill(row + nw, row + db.stride, 0ull);
      db.header_off[i] = hp;
      std::memcpy(db.headers.data() + hp, text + e.header_pos, e.header_len);
      db.headers[hp + e.header_len] = '\0';
      hp += e.header_len + 1;
    }
  });
  db.header_off[n] = part[T];
  clock.lap("gather");
  return true;
}

}  // namespace

void set_ingest_threads(int threads) { g_threads.store(threads); }
unsigned host_threads() { return ingest_threads(); }
static uint64_t ingest_min_bytes() {          // SWARM_B200_INGEST_MIN_BYTES: test hook, lets small inputs take the parallel path
  const char *e = std::getenv("SWARM_B200_INGEST_MIN_BYTES");
  return e ? std::strtoull(e, nullptr, 10) : (1u << 20);
}

std::string db_parse(const char *text, uint64_t size, const DbOptions &opt, AmpliconDb &db) {
  // no more workers than ~64 KiB of text each: a worker needs records to sample (the sort's splitters) and a range to cut
  const unsigned T = static_cast<unsigned>(std::min<uint64_t>(ingest_threads(), std::max<uint64_t>(1, size >> 16)));
  if (T > 1 && size >= ingest_min_bytes() && db_parse_parallel(text, size, opt, db, T)) return "";
  return db_parse_serial(text, size, opt, db);
}

std::string db_read_file(const std::string &path, const DbOptions &opt, AmpliconDb &db) {
  // a regular file is mapped, not copied: the ingest workers then fault their own ranges in, in parallel
  if (path != "-") {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return "Unable to open input data file (" + path + ").\n";          // src/utils/input_output.cc:29-60
    struct stat st;
    if (::fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
      void *map = ::mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE, fd, 0);
      if (map != MAP_FAILED) {
        ::madvise(map, static_cast<size_t>(st.st_size), MADV_WILLNEED);
        std::string e = db_parse(static_cast<const char *>(map), static_cast<uint64_t>(st.st_size), opt, db);
        ::munmap(map, static_cast<size_t>(st.st_size));
        ::close(fd);
        return e;
      }
    }
    ::close(fd);                                                                   // a pipe, an empty file, ...: read it
  }
  std::FILE *f = (path == "-") ? stdin : std::fopen(path.c_str(), "rb");
  if (f == nullptr) return "Unable to open input data file (" + path + ").\n";
  std::vector<char> buf;
  size_t cap = 1u << 24, used = 0;
  buf.resize(cap);
  for (;;) {
    const size_t got = std::fread(buf.data() + used, 1, buf.size() - used, f);
    used += got;
    if (got == 0) break;
    if (used == buf.size()) buf.resize(buf.size() * 2);
  }
  if (f != stdin) std::fclose(f);
  return db_parse(buf.data(), used, opt, db);
}

void append_id(std::string &out, const AmpliconDb &db, uint32_t i, const DbOptions &opt) {   // src/db.cc:946-967
  out.append(db.header(i), db.header_len(i));
  if (opt.append_abundance != 0 && db.ab_start[i] == db.ab_end[i]) {
    if (opt.usearch_abundance) out += ";size=" + std::to_string(db.abundance[i]) + ";";
    else out += "_" + std::to_string(db.abundance[i]);
  }
}

void append_id_noabundance(std::string &out, const AmpliconDb &db, uint32_t i, const DbOptions &opt) {   // :970-998
  const char *h = db.header(i);
  const int32_t hl = static_cast<int32_t>(db.header_len(i)), s = db.ab_start[i], t = db.ab_end[i];
  if (s < t) {
    out.append(h, static_cast<size_t>(s));
    if (opt.usearch_abundance) {
      if (s > 0 && t < hl) out += ';';
      out.append(h + t, static_cast<size_t>(hl - t));
    }
  } else {
    out.append(h, static_cast<size_t>(hl));
  }
}

void append_id_new_abundance(std::string &out, const AmpliconDb &db, uint32_t i, uint64_t abundance, const DbOptions &opt) {  // :1001-1026
  const char *h = db.header(i);
  const int32_t hl = static_cast<int32_t>(db.header_len(i)), s = db.ab_start[i], t = db.ab_end[i];
  out.append(h, static_cast<size_t>(s));
  if (opt.usearch_abundance) {
    if (s > 0) out += ';';
    out += "size=" + std::to_string(abundance) + ";";
    out.append(h + t, static_cast<size_t>(hl - t));
  } else {
    out += "_" + std::to_string(abundance);
  }
}

void append_sequence(std::string &out, const AmpliconDb &db, uint32_t i) {   // src/db.cc:925-943
  static const char sym[4] = {'A', 'C', 'G', 'T'};
  const uint64_t *w = db.seq(i);
  for (uint32_t p = 0; p < db.len[i]; ++p) out += sym[(w[p >> 5] >> ((p & 31) << 1)) & 3];
}

}  // namespace swb
