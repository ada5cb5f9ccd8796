// chain_api.cu - C-ABI entry points for the stages after xcorr_pss (include/lcs_b200.h).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

#include "chain_gpu.hpp"

namespace lcs {

// per-context scratch of the companion kernels, owned by the context (one thread per context, see lcs_b200.h)
ChainScratch& chain_scratch(lcs_ctx* ctx) {
  if (!ctx->chain) ctx->chain = new ChainScratch();
  return *static_cast<ChainScratch*>(ctx->chain);
}
void chain_scratch_release(lcs_ctx* ctx) {
  delete static_cast<ChainScratch*>(ctx->chain);
  ctx->chain = nullptr;
}

static lcs_status upload_c128(lcs_ctx* ctx, const double* capbuf, uint32_t n_cap) {
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));
  LCS_CUDA(ctx, ctx->d_capbuf.ensure((size_t)n_cap * 2));
  LCS_CUDA(ctx, cudaMemcpyAsync(ctx->d_capbuf.p, capbuf, (size_t)n_cap * 16, cudaMemcpyHostToDevice, ctx->streams[0]));
  return LCS_OK;
}

static void to_colmajor(const std::vector<cd>& rowmajor, int n_rows, int n_cols, double* out) {
  cd* o = reinterpret_cast<cd*>(out);
  for (int r = 0; r < n_rows; r++)
    for (int c = 0; c < n_cols; c++) o[(size_t)c * n_rows + r] = rowmajor[(size_t)r * n_cols + c];
}
static std::vector<cd> from_colmajor(const double* in, int n_rows, int n_cols) {
  const cd* i = reinterpret_cast<const cd*>(in);
  std::vector<cd> r((size_t)n_rows * n_cols);
  for (int a = 0; a < n_rows; a++)
    for (int c = 0; c < n_cols; c++) r[(size_t)a * n_cols + c] = i[(size_t)c * n_rows + a];
  return r;
}

// Per-peak stages of CellSearch.cpp:510-558 (sss_detect -> pss_sss_foe -> extract_tfg -> tfoec -> decode_mib) on a
// device-resident capture buffer; cells that fail the SSS or MIB tests are dropped like in the reference.
// Two phases: the device stages run peak by peak on the context's stream; the host stages (tfoec, chan_est, PBCH decoding
// with its 12 tail-biting Viterbi attempts - milliseconds per cell) of all surviving peaks then run on parallel threads.
lcs_status cell_chain_dev(lcs_ctx* ctx, const void* d_cap, int fmt, uint32_t n_cap, const std::vector<lcs_cell>& pk, double fc_req,
                          double fc_prog, double fs_prog, lcs_cell* cells, uint32_t max_cells, uint32_t* n_cells,
                          const int32_t* tracked, uint32_t n_tracked, bool tracker_cycle) {
  const double THRESH2_N_SIGMA = 3;     // CellSearch.cpp:528
  lcs_status rc = LCS_OK;
  if (n_cells) *n_cells = 0;
  if (pk.empty()) return LCS_OK;
  ChainScratch& cs = chain_scratch(ctx);
  struct Pending {
    lcs_cell c;                      // after pss_sss_foe
    std::vector<cd> tfg;
    std::vector<double> ts;
    lcs_cell out{};                  // after decode_mib
  };
  // device stages, each for all peaks at once
  std::vector<lcs_cell> det;
  std::vector<lcs_status> st1;
  rc = dev_sss_detect_batch(ctx, cs, d_cap, fmt, n_cap, pk, THRESH2_N_SIGMA, fc_req, fc_prog, fs_prog, det, st1, nullptr);
  if (rc != LCS_OK) return rc;
  std::vector<lcs_cell> surv;
  for (size_t i = 0; i < pk.size(); i++) {
    if (st1[i] == LCS_ERR_RANGE) continue;   // the reference would index outside the buffer here
    if (det[i].n_id_1 == -1) continue;       // CellSearch.cpp:530-534
    // searcher_thread.cpp:153-174: cells that are being tracked are not examined further
    bool already_tracked = false;
    for (uint32_t k = 0; k < n_tracked; k++) already_tracked |= tracked[k] == det[i].n_id_2 + 3 * det[i].n_id_1;
    if (already_tracked) continue;
    surv.push_back(det[i]);
  }
  std::vector<lcs_cell> foe;
  rc = dev_pss_sss_foe_batch(ctx, cs, d_cap, fmt, n_cap, surv, fc_req, fc_prog, fs_prog, foe);
  if (rc != LCS_OK) return rc;
  std::vector<std::vector<cd>> tfgs;
  std::vector<std::vector<double>> tss;
  std::vector<lcs_status> st3;
  rc = dev_extract_tfg_batch(ctx, cs, d_cap, fmt, n_cap, foe, fc_req, fc_prog, fs_prog, tfgs, tss, st3);
  if (rc != LCS_OK) return rc;
  std::vector<Pending> pend;
  pend.reserve(foe.size());
  for (size_t i = 0; i < foe.size(); i++) {
    if (st3[i] == LCS_ERR_RANGE) continue;
    Pending q;
    q.c = foe[i];
    q.tfg.swap(tfgs[i]);
    q.ts.swap(tss[i]);
    pend.push_back(std::move(q));
  }
  auto host_stage = [&](Pending& q) {
    RsDl rs(q.c.n_id_2 + 3 * q.c.n_id_1, q.c.cp_type);  // CellSearch.cpp:545
    std::vector<cd> tfg_comp(q.tfg.size());
    std::vector<double> ts_comp(q.ts.size());
    lcs_cell o;
    tfoec(q.c, q.tfg.data(), q.ts.data(), (int)q.ts.size(), fc_req, fc_prog, rs, tfg_comp.data(), ts_comp.data(), o);
    decode_mib(o, tfg_comp.data(), (int)q.ts.size(), rs, q.out);
  };
  if (pend.size() <= 1) {
    for (Pending& q : pend) host_stage(q);
  } else {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t n_thr = std::min<size_t>(pend.size(), hw);
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; t++)
      pool.emplace_back([&] {
        for (size_t i = next.fetch_add(1); i < pend.size(); i = next.fetch_add(1)) host_stage(pend[i]);
      });
    for (std::thread& th : pool) th.join();
  }
  // In tracker mode the reference appends every new cell to tracked_cell_list inside its peak loop
  // (searcher_thread.cpp:216-219): a later peak of the same buffer that decodes to an id accepted earlier is skipped.
  uint32_t found = 0;
  std::vector<int> accepted_ids;
  for (Pending& q : pend) {
    if (q.out.n_rb_dl == -1) continue;  // CellSearch.cpp:554-558
    const int id = q.out.n_id_2 + 3 * q.out.n_id_1;
    if (tracker_cycle && std::find(accepted_ids.begin(), accepted_ids.end(), id) != accepted_ids.end()) continue;
    if (found < max_cells && cells) cells[found] = q.out;
    found++;
    accepted_ids.push_back(id);
  }
  if (n_cells) *n_cells = found;
  return LCS_OK;
}

// xcorr_pss + Z_th1 + peak_search (CellSearch.cpp:497-510) on a device-resident capture buffer; only the few values of
// xc_incoherent_single that peak_search reads are fetched.
static lcs_status peaks_dev(lcs_ctx* ctx, const void* d_cap, int fmt, uint32_t n_cap, const double* f_search_set, uint32_t n_f,
                            double fc_req, double fc_prog, double fs_prog, std::vector<lcs_cell>& pk) {
  const uint8_t DS_COMB_ARM = 2;        // CellSearch.cpp:484
  lcs_xcorr_plan* p = nullptr;
  lcs_status rc = get_cached_plan(ctx, n_cap, f_search_set, n_f, DS_COMB_ARM, fc_req, fc_prog, fs_prog, &p);
  if (rc != LCS_OK) return rc;
  cudaStream_t st = ctx->streams[0];
  const size_t n_single = (size_t)3 * n_f * LCS_N_FOLD;
  LCS_CUDA(ctx, ctx->d_single.ensure(n_single));
  LCS_CUDA(ctx, ctx->d_pow.ensure(3 * LCS_N_FOLD));
  LCS_CUDA(ctx, ctx->d_frq.ensure(3 * LCS_N_FOLD));
  LCS_CUDA(ctx, ctx->d_spi.ensure(LCS_N_FOLD));
  rc = plan_run_device(p, d_cap, fmt, 1, ctx->d_single.p, ctx->d_pow.p, ctx->d_frq.p, ctx->d_spi.p, nullptr, st);
  if (rc != LCS_OK) return rc;
  std::vector<double> pw(3 * LCS_N_FOLD), spi(LCS_N_FOLD), z(LCS_N_FOLD);
  std::vector<int32_t> fq(3 * LCS_N_FOLD);
  LCS_CUDA(ctx, cudaMemcpyAsync(pw.data(), ctx->d_pow.p, pw.size() * 8, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaMemcpyAsync(fq.data(), ctx->d_frq.p, fq.size() * 4, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaMemcpyAsync(spi.data(), ctx->d_spi.p, spi.size() * 8, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaStreamSynchronize(st));
  calc_z_th1(spi.data(), LCS_N_FOLD, (uint16_t)p->ps.geom.n_comb_xc, DS_COMB_ARM, z.data());
  // peak_search reads xc_incoherent_single only at (2*arm+1) positions per peak: fetch those on demand.
  cudaError_t fetch_err = cudaSuccess;
  auto single_at = [&](int t, int f, int idx) -> float {
    float v = 0.f;
    cudaError_t e = cudaMemcpy(&v, ctx->d_single.p + ((size_t)t * n_f + f) * LCS_N_FOLD + idx, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) fetch_err = e;
    return v;
  };
  pk.clear();
  peak_search(pw.data(), fq.data(), z.data(), f_search_set, fc_req, fc_prog, single_at, DS_COMB_ARM, pk);
  LCS_CUDA(ctx, fetch_err);
  return LCS_OK;
}

// The chain of CellSearch.cpp:497-558 on a device-resident capture buffer.
static lcs_status cell_search_dev(lcs_ctx* ctx, const void* d_cap, int fmt, uint32_t n_cap, const double* f_search_set,
                                  uint32_t n_f, double fc_req, double fc_prog, double fs_prog, lcs_cell* cells,
                                  uint32_t max_cells, uint32_t* n_cells, lcs_cell* peaks, uint32_t* n_peaks) {
  std::vector<lcs_cell> pk;
  lcs_status rc = peaks_dev(ctx, d_cap, fmt, n_cap, f_search_set, n_f, fc_req, fc_prog, fs_prog, pk);
  if (rc != LCS_OK) return rc;
  if (n_peaks) *n_peaks = (uint32_t)pk.size();
  if (peaks)
    for (size_t i = 0; i < pk.size() && i < max_cells; i++) peaks[i] = pk[i];
  return cell_chain_dev(ctx, d_cap, fmt, n_cap, pk, fc_req, fc_prog, fs_prog, cells, max_cells, n_cells, nullptr, 0);
}

}  // namespace lcs

using namespace lcs;

extern "C" {

lcs_status lcs_calc_z_th1(const double* sp_incoherent, uint32_t n, uint16_t n_comb_xc, uint8_t ds_comb_arm, double* z) {
  if (!sp_incoherent || !z || n_comb_xc == 0) return fail(nullptr, LCS_ERR_ARG, "calc_z_th1: bad argument");
  calc_z_th1(sp_incoherent, n, n_comb_xc, ds_comb_arm, z);
  return LCS_OK;
}

lcs_status lcs_peak_search(const double* pow, const int32_t* frq, const double* z_th1, const double* f_search_set,
                           uint32_t n_f, double fc_requested, double fc_programmed, const float* single_planar,
                           uint8_t ds_comb_arm, lcs_cell* cells, uint32_t max_cells, uint32_t* n_cells) {
  if (!pow || !frq || !z_th1 || !f_search_set || !single_planar || !n_cells || n_f == 0)
    return fail(nullptr, LCS_ERR_ARG, "peak_search: bad argument");
  for (uint32_t i = 0; i < 3 * LCS_N_FOLD; i++)
    if (frq[i] < 0 || (uint32_t)frq[i] >= n_f) return fail(nullptr, LCS_ERR_ARG, "peak_search: frq index out of range");
  std::vector<lcs_cell> v;
  auto at = [&](int t, int f, int idx) { return single_planar[((size_t)t * n_f + f) * LCS_N_FOLD + idx]; };
  peak_search(pow, frq, z_th1, f_search_set, fc_requested, fc_programmed, at, ds_comb_arm, v);
  for (size_t i = 0; i < v.size() && i < max_cells && cells; i++) cells[i] = v[i];
  *n_cells = (uint32_t)v.size();
  return LCS_OK;
}

lcs_status lcs_sss_detect(lcs_ctx* ctx, const lcs_cell* cell, const double* capbuf, uint32_t n_cap, double thresh2_n_sigma,
                          double fc_requested, double fc_programmed, double fs_programmed, lcs_cell* cell_out,
                          double* h1_np, double* h2_np, double* h1_nrm, double* h2_nrm, double* h1_ext, double* h2_ext,
                          double* log_lik_nrm, double* log_lik_ext) {
  if (!ctx || !cell || !capbuf || !cell_out) return fail(ctx, LCS_ERR_ARG, "sss_detect: null argument");
  lcs_status rc = upload_c128(ctx, capbuf, n_cap);
  if (rc != LCS_OK) return rc;
  SssDebugHost d;
  rc = dev_sss_detect(ctx, chain_scratch(ctx), ctx->d_capbuf.p, LCS_IQ_C128, n_cap, *cell, thresh2_n_sigma, fc_requested,
                      fc_programmed, fs_programmed, *cell_out, &d);
  if (rc != LCS_OK) return rc;
  if (h1_np) std::memcpy(h1_np, &d.est[0], 62 * 8);
  if (h2_np) std::memcpy(h2_np, &d.est[62], 62 * 8);
  if (h1_nrm) std::memcpy(h1_nrm, &d.est[124], 62 * 16);
  if (h2_nrm) std::memcpy(h2_nrm, &d.est[248], 62 * 16);
  if (h1_ext) std::memcpy(h1_ext, &d.est[372], 62 * 16);
  if (h2_ext) std::memcpy(h2_ext, &d.est[496], 62 * 16);
  if (log_lik_nrm) std::memcpy(log_lik_nrm, &d.ll[0], 336 * 8);    // mat(168,2) column-major = [col0][col1]
  if (log_lik_ext) std::memcpy(log_lik_ext, &d.ll[336], 336 * 8);
  return LCS_OK;
}

lcs_status lcs_pss_sss_foe(lcs_ctx* ctx, const lcs_cell* cell_in, const double* capbuf, uint32_t n_cap, double fc_requested,
                           double fc_programmed, double fs_programmed, lcs_cell* cell_out) {
  if (!ctx || !cell_in || !capbuf || !cell_out) return fail(ctx, LCS_ERR_ARG, "pss_sss_foe: null argument");
  lcs_status rc = upload_c128(ctx, capbuf, n_cap);
  if (rc != LCS_OK) return rc;
  return dev_pss_sss_foe(ctx, chain_scratch(ctx), ctx->d_capbuf.p, LCS_IQ_C128, n_cap, *cell_in, fc_requested, fc_programmed,
                         fs_programmed, *cell_out);
}

lcs_status lcs_extract_tfg(lcs_ctx* ctx, const lcs_cell* cell, const double* capbuf, uint32_t n_cap, double fc_requested,
                           double fc_programmed, double fs_programmed, double* tfg, double* tfg_timestamp,
                           uint32_t* n_ofdm_out) {
  if (!ctx || !cell || !capbuf || !tfg || !tfg_timestamp) return fail(ctx, LCS_ERR_ARG, "extract_tfg: null argument");
  lcs_status rc = upload_c128(ctx, capbuf, n_cap);
  if (rc != LCS_OK) return rc;
  std::vector<cd> g;
  std::vector<double> ts;
  rc = dev_extract_tfg(ctx, chain_scratch(ctx), ctx->d_capbuf.p, LCS_IQ_C128, n_cap, *cell, fc_requested, fc_programmed,
                       fs_programmed, g, ts);
  if (rc != LCS_OK) return rc;
  to_colmajor(g, (int)ts.size(), 72, tfg);
  std::memcpy(tfg_timestamp, ts.data(), ts.size() * 8);
  if (n_ofdm_out) *n_ofdm_out = (uint32_t)ts.size();
  return LCS_OK;
}

lcs_status lcs_tfoec(lcs_ctx* ctx, const lcs_cell* cell, const double* tfg, const double* tfg_timestamp, uint32_t n_ofdm,
                     double fc_requested, double fc_programmed, double* tfg_comp, double* tfg_comp_timestamp,
                     lcs_cell* cell_out) {
  if (!cell || !tfg || !tfg_timestamp || !tfg_comp || !tfg_comp_timestamp || !cell_out)
    return fail(ctx, LCS_ERR_ARG, "tfoec: null argument");
  if (cell->cp_type != 1 && cell->cp_type != 2) return fail(ctx, LCS_ERR_ARG, "tfoec: cp_type unknown");
  if (cell->n_id_1 < 0 || cell->n_id_2 < 0) return fail(ctx, LCS_ERR_ARG, "tfoec: cell id not set");
  const int n_symb = cell->cp_type == 1 ? 7 : 6;
  if (n_ofdm < (uint32_t)(2 * n_symb)) return fail(ctx, LCS_ERR_ARG, "tfoec: grid too short");
  std::vector<cd> g = from_colmajor(tfg, (int)n_ofdm, 72), gc(g.size());
  RsDl rs(cell->n_id_2 + 3 * cell->n_id_1, cell->cp_type);
  tfoec(*cell, g.data(), tfg_timestamp, (int)n_ofdm, fc_requested, fc_programmed, rs, gc.data(), tfg_comp_timestamp, *cell_out);
  to_colmajor(gc, (int)n_ofdm, 72, tfg_comp);
  return LCS_OK;
}

lcs_status lcs_decode_mib(lcs_ctx* ctx, const lcs_cell* cell, const double* tfg, uint32_t n_ofdm, lcs_cell* cell_out) {
  if (!cell || !tfg || !cell_out) return fail(ctx, LCS_ERR_ARG, "decode_mib: null argument");
  if (cell->cp_type != 1 && cell->cp_type != 2) return fail(ctx, LCS_ERR_ARG, "decode_mib: cp_type unknown");
  if (cell->n_id_1 < 0 || cell->n_id_2 < 0) return fail(ctx, LCS_ERR_ARG, "decode_mib: cell id not set");
  const int n_symb = cell->cp_type == 1 ? 7 : 6;
  if (n_ofdm < (uint32_t)(6 * 20 * n_symb + 2 * n_symb)) return fail(ctx, LCS_ERR_ARG, "decode_mib: grid shorter than 6 frames + 2 slots");
  std::vector<cd> g = from_colmajor(tfg, (int)n_ofdm, 72);
  RsDl rs(cell->n_id_2 + 3 * cell->n_id_1, cell->cp_type);
  decode_mib(*cell, g.data(), (int)n_ofdm, rs, *cell_out);
  return LCS_OK;
}

lcs_status lcs_dedup(const lcs_cell* cells, uint32_t n, lcs_cell* out, uint32_t* n_out) {
  if ((!cells && n) || !out || !n_out) return fail(nullptr, LCS_ERR_ARG, "dedup: null argument");
  std::vector<lcs_cell> fin;
  dedup(cells, n, fin);
  for (size_t i = 0; i < fin.size(); i++) out[i] = fin[i];
  *n_out = (uint32_t)fin.size();
  return LCS_OK;
}

lcs_status lcs_f_search_set(double freq_start, double ppm, double* out, uint32_t* n_f) {
  if (!n_f) return fail(nullptr, LCS_ERR_ARG, "f_search_set: null n_f");
  std::vector<double> f = f_search_set_for(freq_start, ppm);
  if (out) std::memcpy(out, f.data(), f.size() * 8);
  *n_f = (uint32_t)f.size();
  return LCS_OK;
}

lcs_status lcs_cell_search(lcs_ctx* ctx, const double* capbuf, uint32_t n_cap, const double* f_search_set, uint32_t n_f,
                           double fc_requested, double fc_programmed, double fs_programmed, lcs_cell* cells,
                           uint32_t max_cells, uint32_t* n_cells, lcs_cell* peaks, uint32_t* n_peaks) {
  if (!ctx || !capbuf || !f_search_set) return fail(ctx, LCS_ERR_ARG, "cell_search: null argument");
  lcs_status rc = upload_c128(ctx, capbuf, n_cap);
  if (rc != LCS_OK) return rc;
  return cell_search_dev(ctx, ctx->d_capbuf.p, LCS_IQ_C128, n_cap, f_search_set, n_f, fc_requested, fc_programmed,
                         fs_programmed, cells, max_cells, n_cells, peaks, n_peaks);
}

lcs_status lcs_cell_search_cu8(lcs_ctx* ctx, const uint8_t* capbuf_cu8, uint32_t n_cap, const double* f_search_set,
                               uint32_t n_f, double fc_requested, double fc_programmed, double fs_programmed,
                               lcs_cell* cells, uint32_t max_cells, uint32_t* n_cells, lcs_cell* peaks, uint32_t* n_peaks) {
  if (!ctx || !capbuf_cu8 || !f_search_set) return fail(ctx, LCS_ERR_ARG, "cell_search_cu8: null argument");
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));
  LCS_CUDA(ctx, ctx->d_cu8.ensure((size_t)n_cap * 2));
  LCS_CUDA(ctx, cudaMemcpyAsync(ctx->d_cu8.p, capbuf_cu8, (size_t)n_cap * 2, cudaMemcpyHostToDevice, ctx->streams[0]));
  return cell_search_dev(ctx, ctx->d_cu8.p, LCS_IQ_CU8, n_cap, f_search_set, n_f, fc_requested, fc_programmed,
                         fs_programmed, cells, max_cells, n_cells, peaks, n_peaks);
}

// kalibrate (src/LTE-Tracker.cpp:565-741): an initial full search whose only purpose is the oscillator's residual offset.
// The frequency grid is centred on the offset implied by the current correction factor (:586-587); the strongest
// surviving cell after dedup (:703-716) gives freq_superfine and the residual correction factor (:719-726).  The
// reference loops until a cell is found (new data every iteration); here one buffer is examined and *n_cells = 0 reports
// "nothing found, try the next buffer".
lcs_status lcs_kalibrate_cu8(lcs_ctx* ctx, const uint8_t* capbuf_cu8, uint32_t n_cap, double fc_requested, double fc_programmed,
                             double fs_programmed, double ppm, double correction, lcs_cell* best, double* correction_residual,
                             uint32_t* n_cells) {
  if (!ctx || !capbuf_cu8 || !best || !n_cells) return fail(ctx, LCS_ERR_ARG, "kalibrate_cu8: null argument");
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double> f = f_search_set_for(fc_requested, ppm);                          // :586
  for (double& v : f) v = (fc_requested * correction - fc_requested) + v;               // :587
  LCS_CUDA(ctx, ctx->d_cu8.ensure((size_t)n_cap * 2 + 16));
  LCS_CUDA(ctx, cudaMemcpyAsync(ctx->d_cu8.p, capbuf_cu8, (size_t)n_cap * 2, cudaMemcpyHostToDevice, ctx->streams[0]));
  std::vector<lcs_cell> cells(64);
  uint32_t found = 0;
  lcs_status rc = cell_search_dev(ctx, ctx->d_cu8.p, LCS_IQ_CU8, n_cap, f.data(), (uint32_t)f.size(), fc_requested, fc_programmed,
                                  fs_programmed, cells.data(), (uint32_t)cells.size(), &found, nullptr, nullptr);
  if (rc != LCS_OK) return rc;
  std::vector<lcs_cell> fin;
  dedup(cells.data(), std::min<uint32_t>(found, (uint32_t)cells.size()), fin);           // :703-705
  *n_cells = (uint32_t)fin.size();
  lcs_cell_init(best);
  if (fin.empty()) return LCS_OK;
  double bp = -INFINITY;
  for (const lcs_cell& c : fin)                                                         // :709-716
    if (c.pss_pow > bp) { bp = c.pss_pow; *best = c; }
  if (correction_residual) {
    const double true_location = fc_requested;                                          // :720
    const double crystal_freq_actual = fc_programmed - best->freq_superfine;            // :722
    *correction_residual = (true_location / fc_requested * fc_programmed) / crystal_freq_actual;   // :724
  }
  return LCS_OK;
}

// One cycle of the tracker's searcher thread (src/searcher_thread.cpp:95-232) on a capture buffer delivered by the framer.
lcs_status lcs_tracker_search_cu8(lcs_ctx* ctx, const uint8_t* capbuf_cu8, uint32_t n_cap, double frequency_offset,
                                  double fc_requested, double fc_programmed, double fs_programmed, double late,
                                  const int32_t* tracked_n_id_cell, uint32_t n_tracked, lcs_cell* cells, double* frame_timing,
                                  uint32_t max_cells, uint32_t* n_cells) {
  if (!ctx || !capbuf_cu8 || !n_cells || (n_tracked && !tracked_n_id_cell) || (max_cells && (!cells || !frame_timing)))
    return fail(ctx, LCS_ERR_ARG, "tracker_search_cu8: null argument");
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));
  const double f_search_set[1] = {frequency_offset};                                   // searcher_thread.cpp:96-98
  const double k_factor = (fc_requested - frequency_offset) / fc_programmed;
  LCS_CUDA(ctx, ctx->d_cu8.ensure((size_t)n_cap * 2));
  LCS_CUDA(ctx, cudaMemcpyAsync(ctx->d_cu8.p, capbuf_cu8, (size_t)n_cap * 2, cudaMemcpyHostToDevice, ctx->streams[0]));
  std::vector<lcs_cell> pk;
  lcs_status rc = peaks_dev(ctx, ctx->d_cu8.p, LCS_IQ_CU8, n_cap, f_search_set, 1, fc_requested, fc_programmed, fs_programmed, pk);
  if (rc != LCS_OK) return rc;
  uint32_t found = 0;
  rc = cell_chain_dev(ctx, ctx->d_cu8.p, LCS_IQ_CU8, n_cap, pk, fc_requested, fc_programmed, fs_programmed, cells, max_cells, &found,
                      tracked_n_id_cell, n_tracked, true);
  if (rc != LCS_OK) return rc;
  for (uint32_t i = 0; i < found && i < max_cells; i++)                                  // searcher_thread.cpp:214
    frame_timing[i] = cells[i].frame_start * (30720000.0 / 16) / (fs_programmed * k_factor) + late;
  *n_cells = found;
  return LCS_OK;
}

}  // extern "C"
