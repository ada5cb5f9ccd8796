// stream_search_main.cpp - the searcher side of LTE-Tracker on a recorded raw IQ stream (rtl_sdr byte dump):
// producer framing (src/producer_thread.cpp:96-161) + searcher cycles (src/searcher_thread.cpp:83-246) through
// lcs_framer_* and lcs_tracker_search_cu8.  Every new cell is printed with the frame timing the reference would hand
// to its tracker thread and then counts as "tracked".  The tracker threads themselves are out of scope.
//
//   StreamSearch_b200 -f <fc Hz> [-o <frequency offset Hz>] [-c <correction>] [-n <max cycles>] stream.bin
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/lcs_b200.h"

int main(int argc, char** argv) {
  double fc = -1, f_off = 0, correction = 1;
  long max_cycles = -1;
  int c;
  while ((c = getopt(argc, argv, "f:o:c:n:h")) != -1) {
    switch (c) {
      case 'f': fc = strtod(optarg, nullptr); break;
      case 'o': f_off = strtod(optarg, nullptr); break;
      case 'c': correction = strtod(optarg, nullptr); break;
      case 'n': max_cycles = strtol(optarg, nullptr, 10); break;
      default:
        fprintf(stderr, "usage: %s -f <fc Hz> [-o <offset Hz>] [-c <correction>] [-n <max cycles>] stream.bin\n", argv[0]);
        return c == 'h' ? 0 : -1;
    }
  }
  if (fc <= 0 || optind >= argc) {
    fprintf(stderr, "usage: %s -f <fc Hz> [-o <offset Hz>] [-c <correction>] [-n <max cycles>] stream.bin\n", argv[0]);
    return -1;
  }
  FILE* fp = fopen(argv[optind], "rb");
  if (!fp) { perror(argv[optind]); return -1; }
  const double fs_programmed = 1.92e6 * correction, fc_programmed = fc;          // LTE-Tracker.cpp:791,609
  const uint32_t n_cap = 19200 * 8;                                              // LTE-Tracker.cpp:819
  lcs_ctx* ctx = nullptr;
  if (lcs_ctx_create(0, &ctx) != LCS_OK) { fprintf(stderr, "Error: %s\n", lcs_last_error(nullptr)); return -1; }
  lcs_framer* fr = nullptr;
  if (lcs_framer_create(fc, fc_programmed, fs_programmed, n_cap, &fr) != LCS_OK) { fprintf(stderr, "Error: framer\n"); return -1; }
  std::vector<uint8_t> block(2 * 10000);                                         // BLOCK_SIZE, producer_thread.cpp:95
  std::vector<int32_t> tracked;
  long cycles = 0;
  lcs_framer_request(fr);                                                        // searcher_thread.cpp:88
  for (;;) {
    const size_t n = fread(block.data(), 2, 10000, fp);
    if (n == 0) break;
    int ready = 0;
    const uint8_t* cap = nullptr;
    double late = 0;
    if (lcs_framer_push(fr, block.data(), (uint32_t)n, f_off, &ready, &cap, &late) != LCS_OK) { fprintf(stderr, "Error: framer push\n"); return -1; }
    if (!ready) continue;
    lcs_cell cells[16];
    double timing[16];
    uint32_t found = 0;
    if (lcs_tracker_search_cu8(ctx, cap, n_cap, f_off, fc, fc_programmed, fs_programmed, late, tracked.data(), (uint32_t)tracked.size(),
                               cells, timing, 16, &found) != LCS_OK) {
      fprintf(stderr, "Error: %s\n", lcs_last_error(ctx));
      return -1;
    }
    for (uint32_t i = 0; i < found && i < 16; i++) {
      const int id = cells[i].n_id_2 + 3 * cells[i].n_id_1;
      printf("cycle %ld: new cell %d  ports %d  n_rb_dl %d  sfn %d  residual offset %.1f Hz  frame timing %.3f (late %.3f)\n", cycles, id,
             cells[i].n_ports, cells[i].n_rb_dl, cells[i].sfn, cells[i].freq_superfine, timing[i], late);
      tracked.push_back(id);                                                     // searcher_thread.cpp:153-174, 222
    }
    cycles++;
    if (max_cycles >= 0 && cycles >= max_cycles) break;
    lcs_framer_request(fr);
  }
  printf("%ld searcher cycle(s), %zu cell(s) being tracked\n", cycles, tracked.size());
  lcs_framer_destroy(fr);
  lcs_ctx_destroy(ctx);
  fclose(fp);
  return 0;
}
