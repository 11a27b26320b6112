// swarm_b200/host/main.cc — `swarm_b200`: drop-in command line for swarm (d = 0, d = 1, d > 1) on a B200.
//
// Host mirror of /root/reference src/swarm.cc (options :93-124, checks :486-630, dispatch :633-675) and
// src/utils/open_and_close_files.cc:35-93, written from scratch: same options, same validation messages,
// same output bytes (`-o -r -s -i -w -j -u`), same exit code (1 after "\nError: ..." on stderr,
// src/utils/fatal.h:27,38-46).  The clustering itself is the CUDA engine behind include/swarm_b200.h.
// d = 0 (dereplication, src/derep.cc) runs on the engine too; `-u` aligns on the host like the reference
// (uclust.cc, `-t` workers).  `-x`, `-c`, `-y` are accepted and validated but have nothing to configure (no SSE
// dispatch, no Bloom filter to size).
#include "../../include/swarm_b200.h"
#include "../../include/swarm_b200_host.h"

#include <getopt.h>
#include <unistd.h>

#include <array>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace {

[[noreturn]] void fatal(const std::string &msg) {
  std::fprintf(stderr, "\nError: %s\n", msg.c_str());
  std::exit(EXIT_FAILURE);
}

struct Params {
  int64_t append_abundance = 0, boundary = 3, ceiling = 0, differences = 1, gap_ext = 4, gap_open = 12, match = 5,
          mismatch = 4, threads = 1, bloom_bits = 16;
  bool fastidious = false, help = false, ncb = false, mothur = false, version = false, disable_sse3 = false, usearch = false;
  std::string internal_structure, network_file, log, output_file = "-", statistics_file, uclust_file, seeds, input = "-";
  int64_t pen[3] = {18, 24, 13};
};

const option kLong[] = {
    {"append-abundance", required_argument, nullptr, 'a'}, {"boundary", required_argument, nullptr, 'b'},
    {"ceiling", required_argument, nullptr, 'c'}, {"differences", required_argument, nullptr, 'd'},
    {"gap-extension-penalty", required_argument, nullptr, 'e'}, {"fastidious", no_argument, nullptr, 'f'},
    {"gap-opening-penalty", required_argument, nullptr, 'g'}, {"help", no_argument, nullptr, 'h'},
    {"internal-structure", required_argument, nullptr, 'i'}, {"log", required_argument, nullptr, 'l'},
    {"network-file", required_argument, nullptr, 'j'}, {"match-reward", required_argument, nullptr, 'm'},
    {"no-otu-breaking", no_argument, nullptr, 'n'}, {"output-file", required_argument, nullptr, 'o'},
    {"mismatch-penalty", required_argument, nullptr, 'p'}, {"mothur", no_argument, nullptr, 'r'},
    {"statistics-file", required_argument, nullptr, 's'}, {"threads", required_argument, nullptr, 't'},
    {"uclust-file", required_argument, nullptr, 'u'}, {"version", no_argument, nullptr, 'v'},
    {"seeds", required_argument, nullptr, 'w'}, {"disable-sse3", no_argument, nullptr, 'x'},
    {"bloom-bits", required_argument, nullptr, 'y'}, {"usearch-abundance", no_argument, nullptr, 'z'},
    {nullptr, 0, nullptr, 0}};

const char *kHeader =
    "swarm_b200 1.0 — B200 (sm_100a) engine, command-line compatible with Swarm 3.1.6\n"
    "clustering method: Mahe F, Rognes T, Quince C, de Vargas C, Dunthorn M (2014, 2015, 2022), Swarm v1-v3\n\n";

const char *kUsage =
    "Usage: swarm_b200 [OPTIONS] [FASTAFILE]\n\n"
    "General options:\n"
    " -h, --help                          display this help and exit\n"
    " -t, --threads INTEGER               accepted for compatibility (the GPU engine has no thread pool)\n"
    " -v, --version                       display version information and exit\n\n"
    "Clustering options:\n"
    " -d, --differences INTEGER           resolution (1)\n"
    " -n, --no-otu-breaking               never break clusters (not recommended!)\n\n"
    "Fastidious options (only when d = 1):\n"
    " -b, --boundary INTEGER              min mass of large clusters (3)\n"
    " -c, --ceiling INTEGER               accepted for compatibility (no Bloom filter to size)\n"
    " -f, --fastidious                    link nearby low-abundance swarms\n"
    " -y, --bloom-bits INTEGER            accepted for compatibility\n\n"
    "Input/output options:\n"
    " -a, --append-abundance INTEGER      value to use when abundance is missing\n"
    " -i, --internal-structure FILENAME   write internal cluster structure to file\n"
    " -j, --network-file FILENAME         dump sequence network to file\n"
    " -l, --log FILENAME                  log to file, not to stderr\n"
    " -o, --output-file FILENAME          output result to file (stdout)\n"
    " -r, --mothur                        output using mothur-like format\n"
    " -s, --statistics-file FILENAME      dump cluster statistics to file\n"
    " -w, --seeds FILENAME                write cluster representatives to FASTA file\n"
    " -z, --usearch-abundance             abundance annotation in usearch style\n\n"
    "Pairwise alignment advanced options (only when d > 1):\n"
    " -m, --match-reward INTEGER          reward for nucleotide match (5)\n"
    " -p, --mismatch-penalty INTEGER      penalty for nucleotide mismatch (4)\n"
    " -g, --gap-opening-penalty INTEGER   gap open penalty (12)\n"
    " -e, --gap-extension-penalty INTEGER gap extension penalty (4)\n\n";

int64_t args_long(const char *str, const char *option) {   // src/swarm.cc:193-208
  char *end = nullptr;
  const int64_t v = std::strtol(str, &end, 10);
  if (*end != '\0')
    fatal(std::string("Invalid numeric argument for option ") + option +
          ".\n\nFrequent causes are:\n - a missing space between an argument and the next option,\n"
          " - a long option name not starting with a double dash\n   (swarm accepts '--help' or '-h', but not '-help')\n\n"
          "Please see 'swarm --help' for more details.");
  return v;
}

std::FILE *open_out(const std::string &name) {             // src/utils/input_output.cc:44-60: "-" = stdout via dup
  if (name == "-") { const int fd = dup(STDOUT_FILENO); return fd < 0 ? nullptr : fdopen(fd, "w"); }
  return std::fopen(name.c_str(), "w");
}

void write_all(std::FILE *f, char *text, uint64_t len) {
  if (f && text) std::fwrite(text, 1, len, f);
  swbh_free(text);
}

// one line of the reference's progress meter (src/utils/progress.cc:36-80): with -l the finished phase reads "<prompt> 100%";
// on a terminal the reference rewrites the line in place, of which the first and the last state are printed here
void phase_line(std::FILE *logf, bool to_file, const char *prompt) {
  if (to_file) std::fprintf(logf, "%s 100%%\n", prompt);
  else std::fprintf(logf, "%s 0%%  \r%s 100%%\n", prompt, prompt);
}

// |V(x)| summed over the amplicons picked by `want`: 3L substitutions + 3L + 4 insertions + one deletion per homopolymer run
// (src/variants.cc:184-249) — the reference's "Generated ... variants from light swarms" / "Heavy variants" log figures
template <typename Pick>
uint64_t count_microvariants(const swbh_db *db, Pick want) {
  const uint32_t n = swbh_db_count(db), stride = swbh_db_stride_words(db);
  const uint64_t *words = swbh_db_words(db);
  const uint32_t *len = swbh_db_lengths(db);
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (!want(i)) continue;
    const uint64_t *w = words + static_cast<size_t>(i) * stride;
    const uint32_t L = len[i];
    uint64_t runs = 0, prev_top = 0;
    for (uint32_t k = 0; k * 32 < L; ++k) {
      const uint64_t x = w[k] ^ ((w[k] << 2) | prev_top);          // base p against base p - 1
      uint64_t nz = (x | (x >> 1)) & 0x5555555555555555ull;
      const uint32_t in_word = std::min<uint32_t>(32, L - k * 32);
      if (in_word < 32) nz &= (1ull << (2 * in_word)) - 1;
      if (k == 0) nz &= ~1ull;                                     // position 0 opens the first run whatever it holds
      runs += static_cast<uint64_t>(__builtin_popcountll(nz));
      prev_top = w[k] >> 62;
    }
    total += 6ull * L + 4 + (L ? runs + 1 : 0);
  }
  return total;
}

// what the packed database costs on the device, for the out-of-memory message (the layout is fixed-stride)
uint64_t g_db_rows = 0, g_db_stride_words = 0, g_db_nucleotides = 0;

void engine_check(int status) {
  if (status == SWB200_OK) return;
  if (status == SWB200_ENOMEM && g_db_rows) {
    const double gb = static_cast<double>(g_db_rows) * g_db_stride_words * 8 / 1e9, tight = static_cast<double>(g_db_nucleotides) / 4 / 1e9;
    char buf[512];
    std::snprintf(buf, sizeof buf, "GPU engine: %s\nThe packed database is stored with a fixed stride of %" PRIu64 " bytes per sequence (the longest sequence, "
                  "rounded up to 32 nt): %" PRIu64 " sequences need %.1f GB for %.1f GB of nucleotides.%s", swb200_last_error(),
                  g_db_stride_words * 8, g_db_rows, gb, tight,
                  gb > 4 * tight ? "\nA few very long sequences multiply the memory of the whole set; cluster them separately." : "");
    fatal(buf);
  }
  if (status == SWB200_EDUPLICATE)                          // same text as src/algod1.cc:1141-1150
    fatal("some fasta entries have identical sequences.\nSwarm expects dereplicated fasta files.\n"
          "Such files can be produced with swarm or vsearch:\n swarm -d 0 -w derep.fasta -o /dev/null input.fasta\nor\n"
          " vsearch --derep_fulllength input.fasta --sizein --sizeout --output derep.fasta\n");
  fatal(std::string("GPU engine: ") + swb200_last_error());
}

// SWARM_B200_DEVICES="0,1,2,3": the d = 1 clustering as ONE job on several GPUs, all driven by this process,
// one host thread per rank.  The database is SHARDED by rows (no GPU holds it all), every rank hashes its own rows and routes the
// records to the tile owners over peer memory, the join tiles are sharded by hash range and the clustering by amplicon.
// Returns the links of all ranks when `links` is given (-j).
struct RankStatus { int status = SWB200_OK; std::string error; };

template <typename F>
void on_every_rank(uint32_t world, std::vector<RankStatus> &st, F body) {
  std::vector<std::thread> th;
  for (uint32_t r = 0; r < world; ++r)
    th.emplace_back([&, r] {
      if (st[r].status != SWB200_OK) return;
      st[r].status = body(r);
      if (st[r].status != SWB200_OK) st[r].error = swb200_last_error();      // the error text is thread-local
    });
  for (auto &t : th) t.join();
  for (uint32_t r = 0; r < world; ++r)
    if (st[r].status != SWB200_OK) {
      if (st[r].status == SWB200_EDUPLICATE) engine_check(st[r].status);
      fatal("GPU engine (rank " + std::to_string(r) + "): " + st[r].error);
    }
}

void cluster_d1_multi(const std::vector<swb200_ctx *> &ctxs, swbh_db *db, bool ncb, const uint32_t *run_start, uint32_t n_runs,
                      std::vector<uint32_t> &swarm_of, std::vector<uint32_t> &generation, std::vector<uint32_t> &parent,
                      std::vector<uint32_t> *links) {
  const uint32_t world = static_cast<uint32_t>(ctxs.size()), n = swbh_db_count(db), stride = swbh_db_stride_words(db);
  const uint32_t *len = swbh_db_lengths(db);
  const uint32_t lmin = *std::min_element(len, len + n), lmax = *std::max_element(len, len + n);
  const uint32_t per = ((n + world - 1) / world + 1u) & ~1u;      // even: the row staging needs 16-byte aligned shards
  std::vector<RankStatus> st(world);
  on_every_rank(world, st, [&](uint32_t r) {
    swb200_ctx *c = ctxs[r];
    int rc = swb200_set_option(c, "tile_rows", 1);
    if (rc == SWB200_OK) rc = swb200_set_option(c, "job_min_len", lmin);
    if (rc == SWB200_OK) rc = swb200_set_option(c, "job_max_len", lmax);
    if (rc != SWB200_OK) return rc;
    const uint32_t first = std::min<uint64_t>(static_cast<uint64_t>(per) * r, n), count = std::min<uint32_t>(per, n - first);
    return swb200_load_db_rows(c, swbh_db_words(db) + static_cast<size_t>(first) * stride, stride, len + first,
                               swbh_db_abundances(db) + first, n, first, count, run_start, n_runs);
  });
  if (swb200_dist_setup_local(ctxs.data(), world, n, 4) != SWB200_OK) fatal(std::string("GPU engine: ") + swb200_last_error());
  on_every_rank(world, st, [&](uint32_t r) { return swb200_d1_reserve(ctxs[r]); });
  std::vector<std::vector<uint32_t>> out(world * 3);
  std::vector<uint64_t> n_links(world, 0);
  on_every_rank(world, st, [&](uint32_t r) {
    const uint32_t rows = swb200_dist_row_count(n, r, world);
    for (int k = 0; k < 3; ++k) out[r * 3 + k].resize(rows ? rows : 1);
    int rc = swb200_d1_index(ctxs[r]);
    if (rc == SWB200_OK) rc = swb200_d1_network(ctxs[r], ncb ? 1 : 0, &n_links[r]);
    if (rc == SWB200_OK) rc = swb200_d1_cluster_dist(ctxs[r], out[r * 3].data(), out[r * 3 + 1].data(), out[r * 3 + 2].data());
    return rc;
  });
  for (uint32_t r = 0; r < world; ++r) {
    const uint32_t rows = swb200_dist_row_count(n, r, world);
    for (uint32_t i = 0; i < rows; ++i) {
      const uint32_t id = swb200_dist_row_id(r, world, i);
      swarm_of[id] = out[r * 3][i];
      generation[id] = out[r * 3 + 1][i];
      parent[id] = out[r * 3 + 2][i];
    }
  }
  if (links) {
    uint64_t total = 0;
    for (uint64_t m : n_links) total += m;
    links->resize(total * 2 + 2);
    uint64_t at = 0;
    for (uint32_t r = 0; r < world; ++r) {
      if (n_links[r] && swb200_d1_export_links(ctxs[r], links->data() + at * 2) != SWB200_OK) fatal(std::string("GPU engine: ") + swb200_last_error());
      at += n_links[r];
    }
    links->resize(total * 2);
  }
}

}  // namespace

int main(int argc, char **argv) {
  Params P;
  std::array<bool, 26> used{};
  opterr = 1;
  for (;;) {
    int idx = 0;
    const int c = getopt_long(argc, argv, "a:b:c:d:e:fg:hi:j:l:m:no:p:rs:t:u:vw:xy:z", kLong, &idx);
    if (c == -1) break;
    if (c >= 'a' && c <= 'z') {
      if (used[static_cast<size_t>(c - 'a')]) {
        const char *lname = "";
        for (const option *o = kLong; o->name; ++o) if (o->val == c) { lname = o->name; break; }
        fatal(std::string("Option -") + static_cast<char>(c) + " or --" + lname + " specified more than once.");
      }
      used[static_cast<size_t>(c - 'a')] = true;
    }
    switch (c) {
      case 'a': P.append_abundance = args_long(optarg, "-a or --append-abundance"); break;
      case 'b': P.boundary = args_long(optarg, "-b or --boundary"); break;
      case 'c': P.ceiling = args_long(optarg, "-c or --ceiling"); break;
      case 'd': P.differences = args_long(optarg, "-d or --differences"); break;
      case 'e': P.gap_ext = args_long(optarg, "-e or --gap-extension-penalty"); break;
      case 'f': P.fastidious = true; break;
      case 'g': P.gap_open = args_long(optarg, "-g or --gap-opening-penalty"); break;
      case 'h': P.help = true; break;
      case 'i': P.internal_structure = optarg; break;
      case 'j': P.network_file = optarg; break;
      case 'l': P.log = optarg; break;
      case 'm': P.match = args_long(optarg, "-m or --match-reward"); break;
      case 'n': P.ncb = true; break;
      case 'o': P.output_file = optarg; break;
      case 'p': P.mismatch = args_long(optarg, "-p or --mismatch-penalty"); break;
      case 'r': P.mothur = true; break;
      case 's': P.statistics_file = optarg; break;
      case 't': P.threads = args_long(optarg, "-t or --threads"); break;
      case 'u': P.uclust_file = optarg; break;
      case 'v': P.version = true; break;
      case 'w': P.seeds = optarg; break;
      case 'x': P.disable_sse3 = true; break;
      case 'y': P.bloom_bits = args_long(optarg, "-y or --bloom-bits"); break;
      case 'z': P.usearch = true; break;
      default:
        std::fputs(kHeader, stderr);
        std::fputs(kUsage, stderr);
        std::exit(EXIT_FAILURE);
    }
  }
  if (optind < argc) P.input = argv[optind];
  swbh_scoring(P.match, P.mismatch, P.gap_open, P.gap_ext, P.pen);

  // args_check, src/swarm.cc:486-630 (same order, same messages)
  if (P.threads < 1 || P.threads > 512) fatal("Illegal number of threads specified with -t or --threads, must be in the range 1 to 512.");
  // the reference streams `uint8_max` (an unsigned char) into its message, so the upper bound is printed as
  // the single byte 0xFF (src/swarm.cc:513-516 via src/utils/fatal.h); reproduced for byte-identical stderr
  if (P.differences < 0 || P.differences > 255) fatal("Illegal number of differences specified with -d or --differences, must be in the range 0 to \xff.");
  if (P.fastidious && P.differences != 1)
    fatal("Fastidious mode (specified with -f or --fastidious) only works when the resolution (specified with -d or --differences) is 1.");
  if (P.disable_sse3 && P.differences < 2)
    fatal("Option --disable-sse3 or -x has no effect when d < 2 (SSE3 instructions are only used when d > 1).");
  if (!P.fastidious) {
    if (used['b' - 'a']) fatal("Option -b or --boundary specified without -f or --fastidious.");
    if (used['c' - 'a']) fatal("Option -c or --ceiling specified without -f or --fastidious.");
    if (used['y' - 'a']) fatal("Option -y or --bloom-bits specified without -f or --fastidious.");
  }
  if (P.differences < 2) {
    if (used['m' - 'a']) fatal("Option -m or --match-reward specified when d < 2.");
    if (used['p' - 'a']) fatal("Option -p or --mismatch-penalty specified when d < 2.");
    if (used['g' - 'a']) fatal("Option -g or --gap-opening-penalty specified when d < 2.");
    if (used['e' - 'a']) fatal("Option -e or --gap-extension-penalty specified when d < 2.");
  }
  if (P.gap_open < 0) fatal("Illegal gap opening penalty specified with -g or --gap-opening-penalty, must not be negative.");
  if (P.gap_ext < 0) fatal("Illegal gap extension penalty specified with -e or --gap-extension-penalty, must not be negative.");
  if (P.gap_open + P.gap_ext < 1) fatal("Illegal gap penalties specified, the sum of the gap open and the gap extension penalty must be at least 1.");
  if (P.match < 1) fatal("Illegal match reward specified with -m or --match-reward, must be at least 1.");
  if (P.mismatch < 1) fatal("Illegal mismatch penalty specified with -p or --mismatch-penalty, must be at least 1.");
  if (P.boundary < 2) fatal("Illegal boundary specified with -b or --boundary, must be at least 2.");
  if (used['c' - 'a'] && (P.ceiling < 40 || P.ceiling > (1 << 30)))
    fatal("Illegal memory ceiling specified with -c or --ceiling, must be in the range 8 to 1,073,741,824 MB.");
  if (P.bloom_bits < 2 || P.bloom_bits > 64) fatal("Illegal number of Bloom filter bits specified with -y or --bloom-bits, must be in the range 2 to 64.");
  if (used['a' - 'a'] && P.append_abundance < 1) fatal("Illegal abundance value specified with -a or --append-abundance, must be at least 1.");
  if (!P.network_file.empty() && P.differences != 1) fatal("A network file can only written when d = 1.");
  if (P.version) { std::fputs(kHeader, stderr); return EXIT_SUCCESS; }
  if (P.help) { std::fputs(kHeader, stderr); std::fputs(kUsage, stderr); return EXIT_SUCCESS; }
  {
    const int64_t sat16 = std::min<int64_t>(65535 / P.pen[0], (65535 - P.pen[1]) / P.pen[2]);
    if (P.differences > sat16) fatal("Resolution (d) too high for the given scoring system.");
    if (P.pen[0] > 255) fatal("Alignment scoring system yielded a mismatch penalty greater than 255, please use different parameter values.");
  }

  // open_files, src/utils/open_and_close_files.cc:35-93
  std::FILE *out = open_out(P.output_file);
  if (!out) fatal("Unable to open output file for writing.");
  std::FILE *logf = stderr;
  if (!P.log.empty()) { logf = open_out(P.log); if (!logf) fatal("Unable to open log file for writing."); }
  std::FILE *seedsf = nullptr, *statsf = nullptr, *structf = nullptr, *netf = nullptr, *uclustf = nullptr;
  if (!P.seeds.empty() && !(seedsf = open_out(P.seeds))) fatal("Unable to open seeds file for writing.");
  if (!P.statistics_file.empty() && !(statsf = open_out(P.statistics_file))) fatal("Unable to open statistics file for writing.");
  if (!P.uclust_file.empty() && !(uclustf = open_out(P.uclust_file))) fatal("Unable to open uclust file for writing.");
  if (!P.internal_structure.empty() && !(structf = open_out(P.internal_structure))) fatal("Unable to open internal structure file for writing.");
  if (!P.network_file.empty() && !(netf = open_out(P.network_file))) fatal("Unable to open network file for writing.");

  // args_show, src/swarm.cc:211-257
  std::fputs(kHeader, logf);
  std::fprintf(logf, "Database file:     %s\nOutput file:       %s\n", P.input.c_str(), P.output_file.c_str());
  if (!P.statistics_file.empty()) std::fprintf(logf, "Statistics file:   %s\n", P.statistics_file.c_str());
  if (!P.uclust_file.empty()) std::fprintf(logf, "Uclust file:       %s\n", P.uclust_file.c_str());
  if (!P.internal_structure.empty()) std::fprintf(logf, "Int. struct. file  %s\n", P.internal_structure.c_str());
  if (!P.network_file.empty()) std::fprintf(logf, "Network file       %s\n", P.network_file.c_str());
  std::fprintf(logf, "Resolution (d):    %" PRId64 "\nThreads:           %" PRId64 "\n", P.differences, P.threads);
  if (P.differences > 1) {
    std::fprintf(logf, "Scores:            match: %" PRId64 ", mismatch: %" PRId64 "\n", P.match, P.mismatch);
    std::fprintf(logf, "Gap penalties:     opening: %" PRId64 ", extension: %" PRId64 "\n", P.gap_open, P.gap_ext);
    std::fprintf(logf, "Converted costs:   mismatch: %" PRId64 ", gap opening: %" PRId64 ", gap extension: %" PRId64 "\n", P.pen[0], P.pen[1], P.pen[2]);
  }
  std::fprintf(logf, "Break clusters:    %s\n", P.ncb ? "No" : "Yes");
  if (P.fastidious) std::fprintf(logf, "Fastidious:        Yes, with boundary %" PRId64 "\n\n", P.boundary);
  else std::fprintf(logf, "Fastidious:        No\n\n");

  // The CUDA context (driver initialisation, ~0.3-0.5 s) is created on a helper thread while the FASTA is parsed.
  // SWARM_B200_DEVICES = "0,1,2,3": one clustering job on several GPUs (d = 1 without -f; anything else uses the first device).
  std::vector<int> devices;
  if (const char *list = std::getenv("SWARM_B200_DEVICES")) {
    for (const char *p = list; *p;) {
      char *end = nullptr;
      const long v = std::strtol(p, &end, 10);
      if (end == p) break;
      devices.push_back(static_cast<int>(v));
      p = *end == ',' ? end + 1 : end;
    }
  }
  if (devices.empty()) { const char *dev = std::getenv("SWARM_B200_DEVICE"); devices.push_back(dev ? std::atoi(dev) : 0); }
  if (devices.size() > 16 || P.differences != 1 || P.fastidious) devices.resize(1);
  std::vector<swb200_ctx *> ctxs(devices.size(), nullptr);
  swb200_ctx *&ctx = ctxs[0];
  int ctx_status = SWB200_OK;
  std::string ctx_error;
  std::thread ctx_thread([&] {
    for (size_t r = 0; r < devices.size() && ctx_status == SWB200_OK; ++r) {
      ctx_status = swb200_create(&ctxs[r], devices[r]);
      if (ctx_status != SWB200_OK) ctx_error = swb200_last_error();      // the error text is thread-local
    }
  });
  // db_read
  if (used['t' - 'a']) swbh_set_threads(static_cast<int>(P.threads));      // -t bounds the ingest workers; default: all cores
  swbh_db *db = nullptr;
  const int db_status = swbh_db_read_fasta(P.input.c_str(), P.usearch ? 1 : 0, P.append_abundance, P.differences > 1 ? 1 : 0, &db);
  ctx_thread.join();
  if (db_status != 0) fatal(swbh_last_error());
  const uint32_t n = swbh_db_count(db);
  const bool log_to_file = !P.log.empty();
  for (const char *prompt : {"Reading sequences:", "Indexing database:", "Abundance sorting:"}) phase_line(logf, log_to_file, prompt);   // src/db.cc:390,477,675
  std::fprintf(logf, "Database info:     %" PRIu64 " nt in %u sequences, longest %u nt\n", swbh_db_nucleotides(db), n, swbh_db_longest(db));
  g_db_rows = n; g_db_stride_words = swbh_db_stride_words(db); g_db_nucleotides = swbh_db_nucleotides(db);

  swbh_result *res = nullptr;
  char *text = nullptr;
  uint64_t len = 0;
  if (n > 0) {
    if (ctx_status != SWB200_OK) fatal("GPU engine: " + ctx_error);
    {
      const uint16_t *len16 = nullptr; const uint64_t *run_ab = nullptr; const uint32_t *run_start = nullptr;
      const uint32_t runs = swbh_db_compact(db, &len16, &run_ab, &run_start);
      if (ctxs.size() > 1 && runs != 0 && n >= 8192u * ctxs.size() && swbh_db_longest(db) < 8192) {
        // multi-GPU job: every rank uploads its own rows (cluster_d1_multi)
      } else if (runs != 0) engine_check(swb200_load_db_compact(ctx, swbh_db_words(db), swbh_db_stride_words(db), len16, run_ab, run_start, runs, n));
      else engine_check(swb200_load_db(ctx, swbh_db_words(db), swbh_db_stride_words(db), swbh_db_lengths(db), swbh_db_abundances(db), n));
    }
    if (P.differences == 0) {                                  // dereplicate, src/derep.cc:393-418
      std::vector<uint32_t> rep(n), size(n), singles(n);
      std::vector<uint64_t> mass(n);
      uint64_t clusters = 0;
      engine_check(swb200_d0_dereplicate(ctx, rep.data(), mass.data(), size.data(), singles.data(), &clusters));
      for (swb200_ctx *c : ctxs) swb200_destroy(c);
      phase_line(logf, log_to_file, "Dereplicating:    ");                 // src/derep.cc:281, :78, :213-251, :195, :153, :129, :111
      phase_line(logf, log_to_file, "Sorting:          ");
      swbh_derep *dr = nullptr;
      if (swbh_d0_assemble(db, rep.data(), mass.data(), size.data(), singles.data(), &dr) != 0) fatal(swbh_last_error());
      if (swbh_d0_write_swarms(db, dr, P.mothur, P.usearch, P.append_abundance, &text, &len) != 0) fatal(swbh_last_error());
      write_all(out, text, len);
      phase_line(logf, log_to_file, "Writing swarms:   ");
      if (seedsf) { if (swbh_d0_write_seeds(db, dr, P.usearch, &text, &len) != 0) fatal(swbh_last_error()); write_all(seedsf, text, len); phase_line(logf, log_to_file, "Writing seeds:    "); }
      if (uclustf) { if (swbh_d0_write_uclust(db, dr, P.usearch, P.append_abundance, &text, &len) != 0) fatal(swbh_last_error()); write_all(uclustf, text, len); phase_line(logf, log_to_file, "Writing UCLUST:   "); }
      if (structf) { if (swbh_d0_write_structure(db, dr, P.usearch, &text, &len) != 0) fatal(swbh_last_error()); write_all(structf, text, len); phase_line(logf, log_to_file, "Writing structure:"); }
      if (statsf) { if (swbh_d0_write_stats(db, dr, P.usearch, &text, &len) != 0) fatal(swbh_last_error()); write_all(statsf, text, len); phase_line(logf, log_to_file, "Writing stats:    "); }
      std::fprintf(logf, "\nNumber of swarms:  %" PRIu64 "\nLargest swarm:     %u\nHeaviest swarm:    %" PRIu64 "\n", swbh_derep_clusters(dr),
                   swbh_derep_largest(dr), swbh_derep_heaviest(dr));
      swbh_derep_free(dr);
      swbh_db_free(db);
      for (std::FILE *f : {netf, structf, uclustf, statsf, seedsf, out}) if (f) std::fclose(f);
      if (logf != stderr) std::fclose(logf);
      return EXIT_SUCCESS;
    }
    std::vector<uint32_t> swarm_of(n), generation(n), parent(n), extra(n, SWB200_NONE);
    const uint16_t *len16m = nullptr; const uint64_t *run_abm = nullptr; const uint32_t *run_startm = nullptr;
    const uint32_t runs_m = ctxs.size() > 1 ? swbh_db_compact(db, &len16m, &run_abm, &run_startm) : 0;
    if (ctxs.size() > 1 && runs_m != 0 && n >= 8192u * ctxs.size() && swbh_db_longest(db) < 8192) {
      std::vector<uint32_t> links;
      cluster_d1_multi(ctxs, db, P.ncb, run_startm, runs_m, swarm_of, generation, parent, netf ? &links : nullptr);
      phase_line(logf, log_to_file, "Hashing sequences:");
      phase_line(logf, log_to_file, "Building network: ");
      if (netf) {                                              // -j: CSR of the union of the ranks' links, rows ascending (src/algod1.cc:755-788)
        std::vector<uint64_t> row_ptr(static_cast<size_t>(n) + 1, 0);
        const uint64_t m = links.size() / 2;
        for (uint64_t e = 0; e < m; ++e) row_ptr[links[2 * e] + 1]++;
        for (uint32_t i = 0; i < n; ++i) row_ptr[i + 1] += row_ptr[i];
        std::vector<uint32_t> col(m ? m : 1);
        std::vector<uint64_t> cur(row_ptr.begin(), row_ptr.end() - 1);
        for (uint64_t e = 0; e < m; ++e) col[cur[links[2 * e]]++] = links[2 * e + 1];
        for (uint32_t i = 0; i < n; ++i) std::sort(col.begin() + static_cast<int64_t>(row_ptr[i]), col.begin() + static_cast<int64_t>(row_ptr[i + 1]));
        if (swbh_write_network(db, row_ptr.data(), col.data(), P.usearch, P.append_abundance, &text, &len) != 0) fatal(swbh_last_error());
        write_all(netf, text, len);
        phase_line(logf, log_to_file, "Dumping network:  ");
      }
      phase_line(logf, log_to_file, "Clustering:       ");
      if (swbh_d1_assemble(db, swarm_of.data(), generation.data(), parent.data(), nullptr, static_cast<uint64_t>(P.boundary), &res) != 0) fatal(swbh_last_error());
    } else if (P.differences == 1) {
      engine_check(swb200_d1_index(ctx));
      phase_line(logf, log_to_file, "Hashing sequences:");                  // src/algod1.cc:1129, :1162, :759, :1183
      uint64_t links = 0;
      engine_check(swb200_d1_network(ctx, P.ncb ? 1 : 0, &links));
      phase_line(logf, log_to_file, "Building network: ");
      if (netf) {
        std::vector<uint64_t> row_ptr(static_cast<size_t>(n) + 1);
        std::vector<uint32_t> col(links ? links : 1);
        engine_check(swb200_d1_get_network(ctx, row_ptr.data(), col.data()));
        if (swbh_write_network(db, row_ptr.data(), col.data(), P.usearch, P.append_abundance, &text, &len) != 0) fatal(swbh_last_error());
        write_all(netf, text, len);
        phase_line(logf, log_to_file, "Dumping network:  ");
      }
      engine_check(swb200_d1_cluster(ctx, swarm_of.data(), generation.data(), parent.data()));
      phase_line(logf, log_to_file, "Clustering:       ");
      bool grafting = false;
      if (P.fastidious) {
        // the fastidious block of the log, src/algod1.cc:1291-1475
        std::vector<uint32_t> sw_size(n, 0);
        std::vector<uint64_t> sw_mass(n, 0);
        const uint64_t *ab = swbh_db_abundances(db);
        const uint32_t *ln = swbh_db_lengths(db);
        for (uint32_t i = 0; i < n; ++i) { sw_size[swarm_of[i]]++; sw_mass[swarm_of[i]] += ab[i]; }
        uint64_t swarms_before = 0, light_swarms = 0, light_amps = 0, light_nt = 0;
        uint32_t largest_before = 0;
        for (uint32_t i = 0; i < n; ++i)
          if (swarm_of[i] == i) {
            ++swarms_before;
            largest_before = std::max(largest_before, sw_size[i]);
            if (sw_mass[i] < static_cast<uint64_t>(P.boundary)) { ++light_swarms; light_amps += sw_size[i]; }
          }
        for (uint32_t i = 0; i < n; ++i) if (sw_mass[swarm_of[i]] < static_cast<uint64_t>(P.boundary)) light_nt += ln[i];
        std::fprintf(logf, "\nResults before fastidious processing:\nNumber of swarms:  %" PRIu64 "\nLargest swarm:     %u\n\n", swarms_before, largest_before);
        phase_line(logf, log_to_file, "Counting amplicons in heavy and light swarms");
        std::fprintf(logf, "Heavy swarms: %" PRIu64 ", with %" PRIu64 " amplicons\n", swarms_before - light_swarms, n - light_amps);
        std::fprintf(logf, "Light swarms: %" PRIu64 ", with %" PRIu64 " amplicons\n", light_swarms, light_amps);
        std::fprintf(logf, "Total length of amplicons in light swarms: %" PRIu64 "\n", light_nt);
        uint64_t nl = 0, nh = 0;
        engine_check(swb200_d1_fastidious(ctx, static_cast<uint64_t>(P.boundary), extra.data(), &nl, &nh));
        grafting = nl != 0 && nh != 0;
        if (!grafting) {
          std::fprintf(logf, "Only light or heavy swarms found - no need for further analysis.\n");
        } else {
          // the reference sizes a Bloom filter here (bits per entry, k = max(0.4 bits, 1) hash functions, m = 7 x nucleotides x bits,
          // src/algod1.cc:1337-1392); the engine has no filter to size (the graft search is a join), the line is kept for the log's readers
          uint64_t bits = static_cast<uint64_t>(P.bloom_bits);
          uint64_t m = std::max<uint64_t>(light_nt * 7 * bits, 64);
          const unsigned k = std::max(static_cast<unsigned>(0.4 * static_cast<double>(bits)), 1u);
          std::fprintf(logf, "Bloom filter: bits=%" PRIu64 ", m=%" PRIu64 ", k=%u, size=%.1fMB\n", bits, m, k, static_cast<double>(m) / (8.0 * 1048576.0));
          phase_line(logf, log_to_file, "Adding light swarm amplicons to Bloom filter");
          std::fprintf(logf, "Generated %" PRIu64 " variants from light swarms\n",
                       count_microvariants(db, [&](uint32_t i) { return sw_mass[swarm_of[i]] < static_cast<uint64_t>(P.boundary); }));
          phase_line(logf, log_to_file, "Checking heavy swarm amplicons against Bloom filter");
          std::fprintf(logf, "Heavy variants: %" PRIu64 "\n",
                       count_microvariants(db, [&](uint32_t i) { return sw_mass[swarm_of[i]] >= static_cast<uint64_t>(P.boundary); }));
          // the reference counts one candidate per COMMON MICROVARIANT of a (heavy, light) pair; the engine decides pairs (ed <= 2)
          // and reports how many it verified — the one figure of the log that is not the reference's
          uint64_t st[12] = {0};
          swb200_get_stats(ctx, st, 12);
          std::fprintf(logf, "Got %" PRIu64 " graft candidates\n", st[11]);
        }
      }
      if (swbh_d1_assemble(db, swarm_of.data(), generation.data(), parent.data(), grafting ? extra.data() : nullptr,
                           static_cast<uint64_t>(P.boundary), &res) != 0) fatal(swbh_last_error());
      if (grafting) {
        phase_line(logf, log_to_file, "Grafting light swarms on heavy swarms");
        std::fprintf(logf, "Made %" PRIu64 " grafts\n\n", swbh_result_grafts(res));
      }
    } else {
      engine_check(swb200_dn_cluster(ctx, static_cast<uint32_t>(P.differences), P.ncb ? 1 : 0, P.pen, swarm_of.data(), generation.data(),
                                     parent.data(), extra.data()));
      phase_line(logf, log_to_file, "Find qgram vects: ");                  // src/db.cc:831, src/algo.cc:383
      phase_line(logf, log_to_file, "Clustering:       ");
      if (swbh_dn_assemble(db, swarm_of.data(), generation.data(), parent.data(), extra.data(), &res) != 0) fatal(swbh_last_error());
    }
    for (swb200_ctx *c : ctxs) swb200_destroy(c);
    // the writers' progress lines: d = 1 prints one per file (src/algod1.cc:795-1045, order of output_results :1064-1095); d > 1 only
    // around the seeds, where the reference leaves "Collecting seeds:" without its end of line (src/algo.cc:125,163,188)
    const bool d1 = P.differences == 1;
    if (swbh_write_swarms(db, res, P.mothur, P.differences, P.usearch, P.append_abundance, &text, &len) != 0) fatal(swbh_last_error());
    write_all(out, text, len);
    if (d1) phase_line(logf, log_to_file, "Writing swarms:   ");
    if (seedsf) {
      if (swbh_write_seeds(db, res, P.usearch, &text, &len) != 0) fatal(swbh_last_error());
      write_all(seedsf, text, len);
      if (!d1) { std::fprintf(logf, "Collecting seeds:    "); phase_line(logf, log_to_file, "Sorting seeds:    "); }
      phase_line(logf, log_to_file, "Writing seeds:    ");
    }
    if (structf) {
      if ((P.differences == 1 ? swbh_write_structure(db, res, P.usearch, &text, &len) : swbh_dn_write_structure(db, res, P.usearch, &text, &len)) != 0)
        fatal(swbh_last_error());
      write_all(structf, text, len);
      if (d1) phase_line(logf, log_to_file, "Writing structure:");
    }
    if (uclustf) {
      if (swbh_write_uclust(db, res, P.differences, P.pen, P.usearch, P.append_abundance, static_cast<int>(P.threads), &text, &len) != 0)
        fatal(swbh_last_error());
      write_all(uclustf, text, len);
      if (d1) phase_line(logf, log_to_file, "Writing UCLUST:   ");
    }
    if (statsf) {
      if ((P.differences == 1 ? swbh_write_stats(db, res, P.usearch, &text, &len) : swbh_dn_write_stats(db, res, P.usearch, &text, &len)) != 0)
        fatal(swbh_last_error());
      write_all(statsf, text, len);
      if (d1) phase_line(logf, log_to_file, "Writing stats:    ");
    }
    // src/algod1.cc:1484-1487 / src/algo.cc:699-705
    std::fprintf(logf, "\nNumber of swarms:  %" PRIu64 "\nLargest swarm:     %u\nMax generations:   %u\n", swbh_result_swarms(res), swbh_result_largest(res),
                 P.differences == 1 ? swbh_result_maxgen(res) : std::max(1u, swbh_result_maxgen(res)));
    swbh_result_free(res);
  } else {                                                     // empty input: the reference still reports (and the mothur line is written)
    for (swb200_ctx *c : ctxs) if (c) swb200_destroy(c);
    if (P.mothur && P.differences < 2) std::fprintf(out, "swarm_%" PRId64 "\t0\n", P.differences);   // src/algo.cc writes nothing
    std::fprintf(logf, "\nNumber of swarms:  0\nLargest swarm:     0\n%s0\n", P.differences == 0 ? "Heaviest swarm:    " : "Max generations:   ");
  }
  swbh_db_free(db);
  for (std::FILE *f : {netf, structf, uclustf, statsf, seedsf, out}) if (f) std::fclose(f);
  if (logf != stderr) std::fclose(logf);
  return EXIT_SUCCESS;
}
