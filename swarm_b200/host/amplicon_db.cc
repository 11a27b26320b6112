// swarm_b200/host/amplicon_db.cc — FASTA → sorted, 2-bit packed, fixed-stride SoA database.
// Behavioural mirror of /root/reference src/db.cc (citations inline); written from scratch.
#include "amplicon_db.h"

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
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

std::string db_parse(const char *text, uint64_t size, const DbOptions &opt, AmpliconDb &db) {
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

std::string db_read_file(const std::string &path, const DbOptions &opt, AmpliconDb &db) {
  std::FILE *f = (path == "-") ? stdin : std::fopen(path.c_str(), "rb");          // src/utils/input_output.cc:29-60
  if (f == nullptr) return "Unable to open input data file (" + path + ").\n";
  std::vector<char> buf;
  size_t cap = 1u << 24, used = 0;
  if (f != stdin && std::fseek(f, 0, SEEK_END) == 0) {
    const long sz = std::ftell(f);
    if (sz > 0) cap = static_cast<size_t>(sz) + 1;
    std::rewind(f);
  }
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
