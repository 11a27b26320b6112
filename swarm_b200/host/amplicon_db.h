// swarm_b200/host/amplicon_db.h — host-side amplicon database (device-friendly SoA).
//
// Host mirror of the reference's FASTA database layer (/root/reference src/db.cc:432-803 `db_read`,
// :388-413 sort, :852-906 accessors, :925-1026 printers), re-designed for the GPU engine:
//   * sequences are 2-bit packed (A0 C1 G2 T/U3, LSB first — same code as src/db.cc:100-114,561)
//     into a FIXED-STRIDE array `words[n * stride]`, stride = ceil(longest/32) 64-bit words, zero
//     padded, so a tile of consecutive amplicons is one contiguous, 16-byte aligned byte range (TMA
//     bulk copies) and amplicon i starts at i*stride (no offset indirection on the device);
//   * lengths / abundances / header offsets are separate arrays (SoA);
//   * amplicons are stored in the reference's order: abundance descending, then strcmp(header)
//     ascending (src/db.cc:392-406) — index in this order is the amplicon id everywhere downstream.
// Parsing rules, limits and error messages follow SURVEY.md §A.1 / src/db.cc.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace swb {

// std::vector that leaves trivially-constructible elements uninitialised on resize(): the big arrays of the database are
// filled by the ingest workers, a value-initialising resize would first write them once more on one thread
template <typename T>
struct default_init_allocator : std::allocator<T> {
  template <typename U> struct rebind { using other = default_init_allocator<U>; };
  using std::allocator<T>::allocator;
  template <typename U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
  template <typename U, typename... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <typename T> using raw_vector = std::vector<T, default_init_allocator<T>>;

struct DbOptions {
  bool usearch_abundance = false;   // -z  (src/db.cc:214-283)
  int64_t append_abundance = 0;     // -a  (src/db.cc:338-341)
  bool check_duplicate_sequences = false;  // d>1 only (src/db.cc:763-796); d=1 checks on device
};

struct AmpliconDb {
  uint32_t n = 0;
  uint32_t longest = 0;             // nt
  uint32_t longest_header = 0;
  uint32_t stride = 0;              // 64-bit words per amplicon
  uint64_t nucleotides = 0;
  raw_vector<uint64_t> words;       // n * stride
  std::vector<uint32_t> len;        // n
  std::vector<uint64_t> abundance;  // n
  raw_vector<char> headers;         // NUL-terminated, sorted order
  std::vector<uint64_t> header_off; // n + 1
  std::vector<int32_t> ab_start;    // abundance annotation [start, end) inside the header
  std::vector<int32_t> ab_end;

  const char *header(uint32_t i) const { return headers.data() + header_off[i]; }
  uint32_t header_len(uint32_t i) const { return static_cast<uint32_t>(header_off[i + 1] - header_off[i] - 1); }
  const uint64_t *seq(uint32_t i) const { return words.data() + static_cast<uint64_t>(i) * stride; }
};

// Parse a FASTA text held in memory.  Returns "" on success, else the reference's error text
// (without the "\nError: " prefix the caller adds — src/utils/fatal.h:27).
// Large well-formed inputs are parsed by `set_ingest_threads` workers (0 = hardware concurrency, at most 32; 1 = the
// serial parser only); the database and every error message are the same either way.
std::string db_parse(const char *text, uint64_t size, const DbOptions &opt, AmpliconDb &db);
std::string db_parse_serial(const char *text, uint64_t size, const DbOptions &opt, AmpliconDb &db);
void set_ingest_threads(int threads);
unsigned host_threads();             // the worker count set_ingest_threads chose (ingest, output writers)
std::string db_read_file(const std::string &path, const DbOptions &opt, AmpliconDb &db);

// id printers (src/db.cc:946-1026)
void append_id(std::string &out, const AmpliconDb &db, uint32_t i, const DbOptions &opt);
void append_id_noabundance(std::string &out, const AmpliconDb &db, uint32_t i, const DbOptions &opt);
void append_id_new_abundance(std::string &out, const AmpliconDb &db, uint32_t i, uint64_t abundance, const DbOptions &opt);
void append_sequence(std::string &out, const AmpliconDb &db, uint32_t i);   // src/db.cc:925-943

}  // namespace swb
