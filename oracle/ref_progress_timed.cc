// oracle/ref_progress_timed.cc — TEST/BENCH INFRASTRUCTURE, not product code.
//
// Drop-in replacement for the reference's progress meter (interface: /root/reference
// src/utils/progress.h:34-36; behaviour of the original: src/utils/progress.cc:36-80) used only to
// build oracle/_ref/swarm_timed.  It prints the same log text as the original and additionally
// records a steady_clock time per phase; at progress_done() it appends
//     <prompt>\t<seconds>\n
// to the file named by $SWARM_PHASE_TIMES (if set).  Written from scratch for this repo: no
// algorithmic reference file is modified, so cluster outputs are byte-identical to _ref/swarm.
#include "swarm.h"          // struct Parameters (reference header, included where it lies)
#include "utils/opt_log.h"
#include "utils/opt_logfile.h"
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace {
  const char * g_prompt = "";
  uint64_t g_size = 0, g_next = 1, g_chunk = 1;
  std::chrono::steady_clock::time_point g_t0;
}

auto progress_init(const char * prompt, const uint64_t size) -> void
{
  g_prompt = prompt;
  g_size = size;
  g_chunk = size < 200 ? 1 : size / 200;
  g_next = 1;
  if (not opt_log.empty()) { std::fprintf(logfile, "%s", prompt); }
  else { std::fprintf(logfile, "%s %.0f%%", prompt, 0.0); }
  g_t0 = std::chrono::steady_clock::now();
}

auto progress_update(const uint64_t progress) -> void
{
  if (not opt_log.empty()) { return; }
  if (progress < g_next) { return; }
  std::fprintf(logfile, "  \r%s %.0f%%", g_prompt,
               100.0 * static_cast<double>(progress) / static_cast<double>(g_size));
  g_next = progress + g_chunk;
  std::fflush(logfile);
}

auto progress_done(struct Parameters const & parameters) -> void
{
  const double seconds =
    std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count();
  if (not parameters.opt_log.empty()) { std::fprintf(parameters.logfile, " %.0f%%\n", 100.0); }
  else { std::fprintf(parameters.logfile, "  \r%s %.0f%%\n", g_prompt, 100.0); }
  std::fflush(parameters.logfile);
  if (const char * path = std::getenv("SWARM_PHASE_TIMES")) {
    if (std::FILE * f = std::fopen(path, "a")) {
      std::fprintf(f, "%s\t%.6f\n", g_prompt, seconds);
      std::fclose(f);
    }
  }
}
