// stream_host.cpp - host-side framing of a continuous IQ stream into searcher capture buffers, the part of the
// reference's producer thread that feeds the searcher (src/producer_thread.cpp:96-161).  Pure host logic, no CUDA.
//
// The reference keeps a running time stamp in units of LTE samples (FS_LTE/16 = 1.92 MHz) modulo one frame pair
// (19200): every received sample advances it by (FS_LTE/16)/(fs_programmed*k_factor), k_factor =
// (fc_requested - frequency_offset)/fc_programmed.  When the searcher has requested data, the capture starts at the
// first sample whose time stamp is within half a sample of 0 (mod 19200); `late` is that sample's wrapped time stamp
// and is added to frame_start when a detected cell is handed to a tracker (searcher_thread.cpp:214).
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/lcs_b200.h"

struct lcs_framer {
  double fc_requested, fc_programmed, fs_programmed;
  uint32_t n_cap;
  double sample_time = -1;        // producer_thread.cpp:70
  bool request = false, filling = false, ready = false;
  uint32_t idx = 0;
  double late = 0;
  std::vector<uint8_t> capbuf;    // [n_cap][2] raw bytes, handed to the device as they are
};

static inline double wrap(double x, double sm, double lg) {   // macros.h:49 with itpp_ext::matlab_mod
  const double n = lg - sm, v = x - sm;
  return v - n * std::floor(v / n) + sm;
}

extern "C" {

lcs_status lcs_framer_create(double fc_requested, double fc_programmed, double fs_programmed, uint32_t n_cap, lcs_framer** out) {
  if (!out || n_cap == 0 || !(fc_programmed > 0) || !(fs_programmed > 0)) return LCS_ERR_ARG;
  lcs_framer* f = new (std::nothrow) lcs_framer();
  if (!f) return LCS_ERR_STATE;
  f->fc_requested = fc_requested;
  f->fc_programmed = fc_programmed;
  f->fs_programmed = fs_programmed;
  f->n_cap = n_cap;
  f->capbuf.assign((size_t)n_cap * 2, 0);
  *out = f;
  return LCS_OK;
}

void lcs_framer_destroy(lcs_framer* f) { delete f; }

void lcs_framer_request(lcs_framer* f) {
  if (f) { f->request = true; f->ready = false; }
}

double lcs_framer_sample_time(const lcs_framer* f) { return f ? f->sample_time : 0; }

lcs_status lcs_framer_push(lcs_framer* f, const uint8_t* iq, uint32_t n_samples, double frequency_offset, int* ready,
                           const uint8_t** capbuf, double* late) {
  if (!f || (!iq && n_samples) || !ready) return LCS_ERR_ARG;
  const double k_factor = (f->fc_requested - frequency_offset) / f->fc_programmed;      // producer_thread.cpp:99
  const double step = (30720000.0 / 16) / (f->fs_programmed * k_factor);               // :130
  for (uint32_t t = 0; t < n_samples; t++) {
    f->sample_time += step;
    if (f->sample_time > 19200.0) f->sample_time -= 19200.0;                            // :132-133
    const double ts = f->sample_time;
    if (f->request) {
      const double w = wrap(ts - 0, -19200.0 / 2, 19200.0 / 2);
      if (std::fabs(w) < 0.5) {                                                          // :140-146
        f->request = false;
        f->filling = true;
        f->idx = 0;
        f->late = w;
      }
    }
    if (f->filling) {                                                                   // :149-158
      f->capbuf[2 * (size_t)f->idx] = iq[2 * (size_t)t];
      f->capbuf[2 * (size_t)f->idx + 1] = iq[2 * (size_t)t + 1];
      if (++f->idx == f->n_cap) {
        f->filling = false;
        f->ready = true;
      }
    }
  }
  *ready = f->ready ? 1 : 0;
  if (f->ready) {
    if (capbuf) *capbuf = f->capbuf.data();
    if (late) *late = f->late;
  }
  return LCS_OK;
}

}  // extern "C"
