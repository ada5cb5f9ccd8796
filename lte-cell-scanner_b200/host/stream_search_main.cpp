// stream_search_main.cpp - the searcher side of LTE-Tracker on a recorded raw IQ stream (rtl_sdr byte dump):
// producer framing (src/producer_thread.cpp:96-161) + searcher cycles (src/searcher_thread.cpp:83-246) through
// lcs_framer_* and lcs_tracker_search_cu8.  Every new cell is printed with the frame timing the reference would hand
// to its tracker thread and then counts as "tracked".  The tracker threads themselves are out of scope.
//
// Like LTE-Tracker's main (src/LTE-Tracker.cpp:795-798) it first calibrates the oscillator with kalibrate
// (src/LTE-Tracker.cpp:565-741, lcs_kalibrate_cu8) on the first 153600 samples of the stream unless -o gives the offset.
//
//   StreamSearch_b200 -f <fc Hz> [-o <frequency offset Hz>] [-p <ppm>] [-c <correction>] [-n <max cycles>] stream.bin
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/lcs_b200.h"

int main(int argc, char** argv) {
  double fc = -1, f_off = 0, correction = 1, ppm = 120;
  bool have_off = false;
  long max_cycles = -1;
  int c;
  while ((c = getopt(argc, argv, "f:o:c:n:p:h")) != -1) {
    switch (c) {
      case 'f': fc = strtod(optarg, nullptr); break;
      case 'o': f_off = strtod(optarg, nullptr); have_off = true; break;
      case 'p': ppm = strtod(optarg, nullptr); break;
      case 'c': correction = strtod(optarg, nullptr); break;
      case 'n': max_cycles = strtol(optarg, nullptr, 10); break;
      default:
        fprintf(stderr, "usage: %s -f <fc Hz> [-o <offset Hz>] [-p <ppm>] [-c <correction>] [-n <max cycles>] stream.bin\n", argv[0]);
        return c == 'h' ? 0 : -1;
    }
  }
  if (fc <= 0 || optind >= argc) {
    fprintf(stderr, "usage: %s -f <fc Hz> [-o <offset Hz>] [-p <ppm>] [-c <correction>] [-n <max cycles>] stream.bin\n", argv[0]);
    return -1;
  }
  FILE* fp = fopen(argv[optind], "rb");
  if (!fp) { perror(argv[optind]); return -1; }
  const double fs_programmed = 1.92e6 * correction, fc_programmed = fc;          // LTE-Tracker.cpp:791,609
  const uint32_t n_cap = 19200 * 8;                                              // LTE-Tracker.cpp:819
  lcs_ctx* ctx = nullptr;
  if (lcs_ctx_create(0, &ctx) != LCS_OK) { fprintf(stderr, "Error: %s\n", lcs_last_error(nullptr)); return -1; }
  if (!have_off) {
    // kalibrate: "similar to running CellSearch with only one center frequency; all information is discarded except for
    // the frequency offset" (LTE-Tracker.cpp:793-798)
    std::vector<uint8_t> first((size_t)n_cap * 2);
    if (fread(first.data(), 2, n_cap, fp) != n_cap) { fprintf(stderr, "Error: not enough data in file!\n"); return -1; }
    rewind(fp);
    lcs_cell best;
    double resid = 1;
    uint32_t n_found = 0;
    printf("Calibrating local oscillator.\n");
    if (lcs_kalibrate_cu8(ctx, first.data(), n_cap, fc, fc_programmed, fs_programmed, ppm, correction, &best, &resid, &n_found) != LCS_OK) {
      fprintf(stderr, "Error: %s\n", lcs_last_error(ctx));
      return -1;
    }
    if (!n_found) { printf("Calibration failed (no cells detected).\n"); return 1; }
    printf("Calibration succeeded!\n   Residual frequency offset: %g Hz\n   New correction factor: %.20g\n", best.freq_superfine, resid);
    f_off = best.freq_superfine;                                                 // global_thread_data.frequency_offset(initial_freq_offset)
  }
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
