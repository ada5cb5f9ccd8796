// lcs_oracle_c.cpp - extern "C" surface of the CPU oracle for ctypes (tests/, bench.py's
// cpu_baseline / --impl reference leg, __graft_entry__.smoke()).  TEST INFRASTRUCTURE ONLY.
#include <cstring>

#include "lcs_oracle.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace lcso;

extern "C" {

int lcso_cell_sizeof() { return (int)sizeof(Cell); }
void lcso_cell_init(Cell* c) { *c = make_cell(); }
int lcso_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void lcso_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

double lcso_chi2cdf_inv(double p, double k) { return chi2cdf_inv(p, k); }
void lcso_lte_pn(uint32_t c_init, uint32_t len, uint8_t* out) {
  auto v = lte_pn(c_init, len);
  memcpy(out, v.data(), len);
}
void lcso_pss_td(int t, double* out274) {
  auto v = pss_td_calc(t);
  memcpy(out274, v.data(), 137 * 16);
}
void lcso_pss_fd(int t, double* out124) {
  auto v = pss_fd_calc(t);
  memcpy(out124, v.data(), 62 * 16);
}
void lcso_sss_fd(int n_id_1, int n_id_2, int slot, int32_t* out62) {
  auto v = sss_fd_calc(n_id_1, n_id_2, slot);
  for (int i = 0; i < 62; i++) out62[i] = v[i];
}
// rs[20*n_symb][12] complex (NaN rows where no RS), shift[20*n_symb][4]
void lcso_rs_dl(int n_id_cell, int cp_type, double* rs, double* shift) {
  RS_DL r(n_id_cell, 6, cp_type);
  for (size_t i = 0; i < r.table.size(); i++)
    for (int k = 0; k < 12; k++) {
      rs[(i * 12 + k) * 2] = r.table[i].empty() ? NAN : r.table[i][k].real();
      rs[(i * 12 + k) * 2 + 1] = r.table[i].empty() ? NAN : r.table[i][k].imag();
    }
  memcpy(shift, r.shift_table.data(), r.shift_table.size() * 8);
}
void lcso_conv_encode(const uint8_t* c, int n, uint8_t* d) {
  auto v = lte_conv_encode(std::vector<uint8_t>(c, c + n));
  memcpy(d, v.data(), 3 * n);
}
void lcso_conv_decode(const double* d_est, int n_c, uint8_t* c) {
  auto v = lte_conv_decode(std::vector<double>(d_est, d_est + 3 * n_c), n_c);
  memcpy(c, v.data(), n_c);
}
void lcso_crc16(const uint8_t* a, int n, uint8_t* p16) {
  auto v = lte_calc_crc16(std::vector<uint8_t>(a, a + n));
  memcpy(p16, v.data(), 16);
}
void lcso_deratematch(const double* e, int n_e, int n_c, double* d) {
  auto v = lte_conv_deratematch(std::vector<double>(e, e + n_e), n_c);
  memcpy(d, v.data(), 3 * n_c * 8);
}
void lcso_f_search_set(double freq_start, double ppm, double* out, int* n) {
  auto v = f_search_set_for(freq_start, ppm);
  *n = (int)v.size();
  if (out) memcpy(out, v.data(), v.size() * 8);
}

// xcorr_pss (searcher.h:22-41).  Layouts: pow/frq [t][idx]; single/incoherent [t][idx][f] as double;
// xc [t][k][f] interleaved re/im doubles (nullable); sp [n_comb_sp*9600] (nullable).
int lcso_xcorr_pss(const double* capbuf, uint32_t n_cap, const double* f_search_set, int n_f, int ds_comb_arm,
                   double fc_requested, double fc_programmed, double fs_programmed, uint32_t flags, double* pow,
                   int32_t* frq, double* single, double* incoherent, double* sp_incoherent, double* xc, double* sp,
                   uint16_t* n_comb_xc, uint16_t* n_comb_sp) {
  XcorrOut o;
  xcorr_pss((const cd*)capbuf, n_cap, f_search_set, n_f, ds_comb_arm, fc_requested, fc_programmed, fs_programmed, flags,
            xc != nullptr, o);
  if (pow) memcpy(pow, o.pow.data(), o.pow.size() * 8);
  if (frq) memcpy(frq, o.frq.data(), o.frq.size() * 4);
  if (single) memcpy(single, o.single.data(), o.single.size() * 8);
  if (incoherent) memcpy(incoherent, o.incoherent.data(), o.incoherent.size() * 8);
  if (sp_incoherent) memcpy(sp_incoherent, o.sp_incoherent.data(), 9600 * 8);
  if (xc) memcpy(xc, o.xc.data(), o.xc.size() * 16);
  if (sp) memcpy(sp, o.sp.data(), o.sp.size() * 8);
  if (n_comb_xc) *n_comb_xc = o.n_comb_xc;
  if (n_comb_sp) *n_comb_sp = o.n_comb_sp;
  return 0;
}

void lcso_calc_Z_th1(const double* sp_incoherent, int n_comb_xc, int ds_comb_arm, double* Z) {
  auto v = calc_Z_th1(std::vector<double>(sp_incoherent, sp_incoherent + 9600), n_comb_xc, ds_comb_arm);
  memcpy(Z, v.data(), 9600 * 8);
}

int lcso_peak_search(const double* pow, const int32_t* frq, const double* Z_th1, const double* f_search_set, int n_f,
                     double fc_requested, double fc_programmed, const double* single, int ds_comb_arm, Cell* cells,
                     int max_cells) {
  std::vector<Cell> v;
  peak_search(pow, frq, Z_th1, f_search_set, n_f, fc_requested, fc_programmed, single, ds_comb_arm, v);
  for (int i = 0; i < (int)v.size() && i < max_cells; i++) cells[i] = v[i];
  return (int)v.size();
}

// dbg (nullable): [h1_np 62][h2_np 62][h1_nrm 124][h2_nrm 124][h1_ext 124][h2_ext 124][ll_nrm 336][ll_ext 336]
void lcso_sss_detect(const Cell* cell, const double* capbuf, uint32_t n_cap, double thresh2_n_sigma, double fc_requested,
                     double fc_programmed, double fs_programmed, uint32_t flags, Cell* out, double* dbg) {
  SssDebug d;
  *out = sss_detect(*cell, (const cd*)capbuf, n_cap, thresh2_n_sigma, fc_requested, fc_programmed, fs_programmed, flags, d);
  if (dbg) {
    memcpy(dbg, d.h1_np.data(), 62 * 8);
    memcpy(dbg + 62, d.h2_np.data(), 62 * 8);
    memcpy(dbg + 124, d.h1_nrm.data(), 62 * 16);
    memcpy(dbg + 248, d.h2_nrm.data(), 62 * 16);
    memcpy(dbg + 372, d.h1_ext.data(), 62 * 16);
    memcpy(dbg + 496, d.h2_ext.data(), 62 * 16);
    memcpy(dbg + 620, d.log_lik_nrm.data(), 336 * 8);
    memcpy(dbg + 956, d.log_lik_ext.data(), 336 * 8);
  }
}
void lcso_pss_sss_foe(const Cell* cell, const double* capbuf, uint32_t n_cap, double fc_requested, double fc_programmed,
                      double fs_programmed, uint32_t flags, Cell* out) {
  *out = pss_sss_foe(*cell, (const cd*)capbuf, n_cap, fc_requested, fc_programmed, fs_programmed, flags);
}
// tfg: [n_ofdm][72] interleaved; returns n_ofdm
int lcso_extract_tfg(const Cell* cell, const double* capbuf, uint32_t n_cap, double fc_requested, double fc_programmed,
                     double fs_programmed, uint32_t flags, double* tfg, double* ts) {
  std::vector<cd> g;
  std::vector<double> t;
  extract_tfg(*cell, (const cd*)capbuf, n_cap, fc_requested, fc_programmed, fs_programmed, flags, g, t);
  memcpy(tfg, g.data(), g.size() * 16);
  memcpy(ts, t.data(), t.size() * 8);
  return (int)t.size();
}
void lcso_tfoec(const Cell* cell, const double* tfg, const double* ts, int n_ofdm, double fc_requested,
                double fc_programmed, uint32_t flags, double* tfg_comp, double* ts_comp, Cell* out) {
  std::vector<cd> g((const cd*)tfg, (const cd*)tfg + (size_t)n_ofdm * 72), gc;
  std::vector<double> t(ts, ts + n_ofdm), tc;
  RS_DL rs(cell->n_id_cell(), 6, cell->cp_type);
  *out = tfoec(*cell, g, t, fc_requested, fc_programmed, rs, flags, gc, tc);
  memcpy(tfg_comp, gc.data(), gc.size() * 16);
  memcpy(ts_comp, tc.data(), tc.size() * 8);
}
// ce: [n_ofdm][72] interleaved
void lcso_chan_est(const Cell* cell, const double* tfg, int n_ofdm, int port, double* ce, double* np) {
  std::vector<cd> g((const cd*)tfg, (const cd*)tfg + (size_t)n_ofdm * 72), c;
  RS_DL rs(cell->n_id_cell(), 6, cell->cp_type);
  chan_est(*cell, rs, g, n_ofdm, port, c, *np);
  memcpy(ce, c.data(), c.size() * 16);
}
// c_est40 (nullable): the 40 decoded bits of the successful attempt; returns frame_timing_guess or -1
int lcso_decode_mib(const Cell* cell, const double* tfg, int n_ofdm, Cell* out, uint8_t* c_est40, double* np_v4) {
  std::vector<cd> g((const cd*)tfg, (const cd*)tfg + (size_t)n_ofdm * 72);
  RS_DL rs(cell->n_id_cell(), 6, cell->cp_type);
  MibDebug d;
  *out = decode_mib(*cell, g, n_ofdm, rs, &d);
  if (c_est40 && d.c_est.size() == 40) memcpy(c_est40, d.c_est.data(), 40);
  if (np_v4) memcpy(np_v4, d.np_v, 32);
  return d.frame_timing_guess;
}
int lcso_dedup(const Cell* in, int n, Cell* out) {
  std::vector<std::vector<Cell>> det(1, std::vector<Cell>(in, in + n));
  std::vector<Cell> fin;
  dedup(det, fin);
  for (size_t i = 0; i < fin.size(); i++) out[i] = fin[i];
  return (int)fin.size();
}
// One centre frequency of CellSearch's main loop.  peaks (nullable) receives the peak_search list.
int lcso_cell_search_one(const double* capbuf, uint32_t n_cap, const double* f_search_set, int n_f, double fc_requested,
                         double fc_programmed, double fs_programmed, uint32_t flags, Cell* cells, int max_cells,
                         Cell* peaks, int* n_peaks) {
  std::vector<Cell> v, pk;
  cell_search_one((const cd*)capbuf, n_cap, f_search_set, n_f, fc_requested, fc_programmed, fs_programmed, flags, v, &pk);
  for (int i = 0; i < (int)v.size() && i < max_cells; i++) cells[i] = v[i];
  if (peaks)
    for (int i = 0; i < (int)pk.size() && i < max_cells; i++) peaks[i] = pk[i];
  if (n_peaks) *n_peaks = (int)pk.size();
  return (int)v.size();
}

}  // extern "C"
