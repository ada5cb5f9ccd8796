// search_batch.cu - threshold + peak_search on the device and the batched search entry points built on it
// (no host round trip between xcorr_pss and the per-peak stages of a sweep).
//
//   peak_search_kernel          src/searcher.cpp:422-510 with Z_th1 of src/CellSearch.cpp:500-503 folded in
//   lcs_xcorr_peaks_batch_host  xcorr_pss + threshold + peak_search for a batch of host capture buffers
//   lcs_cell_search_batch_cu8   the whole chain of CellSearch.cpp:497-558 per buffer of a batch
#include <cmath>
#include <cstring>
#include <vector>

#include "chain_gpu.hpp"
#include "chain_host.hpp"
#include "lcs_ctx.hpp"

namespace lcs {

struct DevPeak {
  double pss_pow;
  int32_t ind;      // refined index (-1: the reference's uint16 wrap, searcher.cpp:459)
  int32_t fi;       // index into f_search_set
  int32_t row;      // n_id_2
  int32_t col;      // peak column before refinement
};

constexpr int PK_THREADS = 512;

// (value, flat index) arg-max with the reference's tie rule: first maximum of each row, rows compared with a strict
// '>' in order 0,1,2 (searcher.cpp:441-445)  ==  largest value, smallest flat index row*9600+col among equals.
__device__ __forceinline__ void argmax_combine(double& v, int& i, double v2, int i2) {
  if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
}

// One CTA per capture buffer.  `work` ([batch][3][9600] double scratch) is written only once a peak has been found:
// buffers without a PSS above threshold cost a single pass over `pow`.
__global__ void __launch_bounds__(PK_THREADS) peak_search_kernel(const double* __restrict__ pow_all, const int32_t* __restrict__ frq_all,
                                                                 const double* __restrict__ spi_all, const float* __restrict__ single_all,
                                                                 double* __restrict__ work_all, DevPeak* __restrict__ peaks_all,
                                                                 int32_t* __restrict__ npeaks_all, uint32_t n_f /* stride */, double r_th1,
                                                                 double rx_cutoff, double n_comb, double box, int arm, double cancel,
                                                                 int max_peaks) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = 3 * LCS_N_FOLD;
  const double* pw = pow_all + (size_t)b * N;
  const int32_t* frq = frq_all + (size_t)b * N;
  const double* spi = spi_all + (size_t)b * LCS_N_FOLD;
  const float* single = single_all + (size_t)b * 3 * n_f * LCS_N_FOLD;
  double* work = work_all + (size_t)b * N;
  __shared__ double s_v[PK_THREADS / 32];
  __shared__ int s_i[PK_THREADS / 32];
  __shared__ double s_best;
  __shared__ int s_besti, s_stop;
  const double* src = pw;
  int n_found = 0;
  for (;;) {
    double v = -INFINITY;
    int vi = 0x7fffffff;
    for (int i = tid; i < N; i += PK_THREADS) argmax_combine(v, vi, src[i], i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, vi, o);
      argmax_combine(v, vi, v2, i2);
    }
    if (lane == 0) { s_v[warp] = v; s_i[warp] = vi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < PK_THREADS / 32; w++) argmax_combine(v, vi, s_v[w], s_i[w]);
      const int row = vi / LCS_N_FOLD, col = vi - row * LCS_N_FOLD;
      // Z_th1 (CellSearch.cpp:500-503), same operation order as the host code
      const double z = r_th1 * spi[col] / rx_cutoff / 137 / 2 / n_comb / box;
      int stop = (v < z) || !(v > 0);                                    // searcher.cpp:446 (+ all-zero guard as on the host)
      if (!stop && n_found >= max_peaks) stop = 2;                        // overflow: the host redoes this buffer
      if (!stop) {
        const int fi = frq[vi];
        int ind = -1;                                                     // searcher.cpp:457-465 incl. the uint16 wrap
        if (col >= arm) {
          float bp = -INFINITY;
          for (int t = col - arm; t <= col + arm; t++) {
            const int tw = t % LCS_N_FOLD;
            const float sv = single[((size_t)row * n_f + fi) * LCS_N_FOLD + tw];
            if (sv > bp) { bp = sv; ind = tw; }
          }
        }
        DevPeak pk;
        pk.pss_pow = v; pk.ind = ind; pk.fi = fi; pk.row = row; pk.col = col;
        peaks_all[(size_t)b * max_peaks + n_found] = pk;
      }
      s_best = v; s_besti = vi; s_stop = stop;
    }
    __syncthreads();
    if (s_stop) {
      if (tid == 0) npeaks_all[b] = s_stop == 2 ? max_peaks + 1 : n_found;
      return;
    }
    n_found++;
    const double best = s_best;
    const int row = s_besti / LCS_N_FOLD, col = s_besti - row * LCS_N_FOLD;
    const double th = best * cancel;                                      // searcher.cpp:501
    // no second peak of the same PSS within +-274 samples (:481-484); drop everything 12 dB below this peak (:501-508)
    for (int i = tid; i < N; i += PK_THREADS) {
      double x = src[i];
      const int r = i / LCS_N_FOLD, c = i - r * LCS_N_FOLD;
      if (r == row) {
        int d = c - col;
        if (d < 0) d = -d;
        if (d > LCS_N_FOLD / 2) d = LCS_N_FOLD - d;                       // circular distance
        if (d <= 274) x = 0;
      }
      if (x < th) x = 0;
      work[i] = x;
    }
    src = work;
    __syncthreads();
  }
}

static lcs_status launch_peak_search(lcs_ctx* ctx, const XcorrGeom& g, uint32_t nb, const double* d_pow, const int32_t* d_frq,
                                     const double* d_spi, const float* d_single, double* d_work, DevPeak* d_peaks, int32_t* d_npeaks,
                                     int max_peaks, cudaStream_t st) {
  const double r_th1 = chi2cdf_inv(1 - std::pow(10.0, -12.0), 2.0 * g.n_comb_xc * (2 * g.ds_comb_arm + 1));
  const double rx_cutoff = (6 * 12 * 15e3 / 2 + 4 * 15e3) / ((30720000.0 / 16) / 2);
  const double cancel = std::pow(10.0, -12.0 / 10.0);
  peak_search_kernel<<<nb, PK_THREADS, 0, st>>>(d_pow, d_frq, d_spi, d_single, d_work, d_peaks, d_npeaks, g.n_f_stride, r_th1, rx_cutoff,
                                                (double)g.n_comb_xc, (double)(2 * g.ds_comb_arm + 1), (int)g.ds_comb_arm, cancel,
                                                max_peaks);
  ctx->launches++;
  LCS_CUDA(ctx, cudaGetLastError());
  return LCS_OK;
}

constexpr int SEARCH_MAX_PEAKS = 32;
constexpr uint32_t SEARCH_CHUNK = 64;

// Shared driver: xcorr_pss + threshold + peak_search for `batch` host capture buffers in chunks that alternate between the
// context's two streams; `per_buffer(buffer index, device pointer of the buffer's IQ bytes, its PSS peaks)` runs on the
// host after the chunk's kernels finished while the next chunk is already in flight on the other stream.
// d_buf_plan == NULL: every buffer is searched with plan 0 of `ps`; otherwise buffer b uses plan d_buf_plan[b] and
// h_buf_plan[b] names the same plan on the host.
template <class F>
static lcs_status search_chunks(lcs_ctx* ctx, PlanSet& ps, int kernel, lcs_xcorr_plan::HostBatchBufs (&hbs)[lcs_ctx::N_STREAMS], const void* h_iq,
                                int iq_format, uint32_t batch, uint32_t chunk, const uint32_t* d_buf_plan, const uint32_t* h_buf_plan,
                                F&& per_buffer) {
  const XcorrGeom& g = ps.geom;
  const size_t samp_bytes = iq_format == LCS_IQ_CU8 ? 2 : (iq_format == LCS_IQ_CF32 ? 8 : (iq_format == LCS_IQ_C128 ? 16 : 0));
  if (!samp_bytes) return fail(ctx, LCS_ERR_ARG, "search_batch: bad iq_format");
  if (batch == 0) return LCS_OK;
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));
  chunk = std::max<uint32_t>(1, std::min<uint32_t>(chunk, batch));
  const size_t n_single = (size_t)3 * g.n_f_stride * LCS_N_FOLD;
  // 16-byte aligned buffer stride inside a chunk is not required (the correlator aligns absolute addresses), only the base
  constexpr int NS = lcs_ctx::N_STREAMS;
  for (int s = 0; s < NS; s++) {
    auto& hb = hbs[s];
    LCS_CUDA(ctx, hb.iq.ensure((size_t)chunk * g.n_cap * samp_bytes + 16));
    LCS_CUDA(ctx, hb.single.ensure(chunk * n_single));
    LCS_CUDA(ctx, hb.pow.ensure((size_t)chunk * 3 * LCS_N_FOLD));
    LCS_CUDA(ctx, hb.frq.ensure((size_t)chunk * 3 * LCS_N_FOLD));
    LCS_CUDA(ctx, hb.spi.ensure((size_t)chunk * LCS_N_FOLD));
    LCS_CUDA(ctx, hb.work.ensure((size_t)chunk * 3 * LCS_N_FOLD));
    LCS_CUDA(ctx, hb.peaks.ensure((size_t)chunk * SEARCH_MAX_PEAKS * sizeof(DevPeak)));
    LCS_CUDA(ctx, hb.npeaks.ensure(chunk));
    // page-locked: with pageable memory cudaMemcpyAsync would block the host until the chunk's kernels are done and the
    // next chunk could not be queued behind it
    LCS_CUDA(ctx, hb.h_peaks.ensure((size_t)chunk * SEARCH_MAX_PEAKS * sizeof(DevPeak)));
    LCS_CUDA(ctx, hb.h_npeaks.ensure(chunk));
  }
  const DevPeak* h_peaks[NS];
  const int32_t* h_np[NS];
  for (int s = 0; s < NS; s++) {
    h_peaks[s] = reinterpret_cast<const DevPeak*>(hbs[s].h_peaks.p);
    h_np[s] = hbs[s].h_npeaks.p;
  }
  auto issue = [&](uint32_t b0, int s) -> lcs_status {
    const uint32_t nb = std::min(chunk, batch - b0);
    cudaStream_t st = ctx->streams[s];
    auto& hb = hbs[s];
    LCS_CUDA(ctx, cudaMemcpyAsync(hb.iq.p, (const char*)h_iq + (size_t)b0 * g.n_cap * samp_bytes, (size_t)nb * g.n_cap * samp_bytes,
                                  cudaMemcpyHostToDevice, st));
    lcs_status rc = planset_run(ps, kernel, hb.iq.p, iq_format, nb, d_buf_plan ? d_buf_plan + b0 : nullptr, hb.single.p, hb.pow.p,
                                hb.frq.p, hb.spi.p, nullptr, st);
    if (rc != LCS_OK) return rc;
    rc = launch_peak_search(ctx, g, nb, hb.pow.p, hb.frq.p, hb.spi.p, hb.single.p, hb.work.p, reinterpret_cast<DevPeak*>(hb.peaks.p),
                            hb.npeaks.p, SEARCH_MAX_PEAKS, st);
    if (rc != LCS_OK) return rc;
    LCS_CUDA(ctx, cudaMemcpyAsync(hb.h_npeaks.p, hb.npeaks.p, nb * 4, cudaMemcpyDeviceToHost, st));
    LCS_CUDA(ctx, cudaMemcpyAsync(hb.h_peaks.p, hb.peaks.p, (size_t)nb * SEARCH_MAX_PEAKS * sizeof(DevPeak), cudaMemcpyDeviceToHost, st));
    return LCS_OK;
  };
  auto finish = [&](uint32_t b0, int s) -> lcs_status {
    const uint32_t nb = std::min(chunk, batch - b0);
    auto& hb = hbs[s];
    LCS_CUDA(ctx, cudaStreamSynchronize(ctx->streams[s]));
    std::vector<lcs_cell> pk;
    for (uint32_t i = 0; i < nb; i++) {
      const PlanCfg& cfg = ps.cfg[h_buf_plan ? h_buf_plan[b0 + i] : 0];
      pk.clear();
      if (h_np[s][i] > SEARCH_MAX_PEAKS) {
        // more peaks than the device list holds: redo this buffer's peak_search on the host (src/searcher.cpp:422-510)
        std::vector<double> pw(3 * LCS_N_FOLD), spi(LCS_N_FOLD), z(LCS_N_FOLD);
        std::vector<int32_t> fq(3 * LCS_N_FOLD);
        std::vector<float> sg(n_single);
        LCS_CUDA(ctx, cudaMemcpy(pw.data(), hb.pow.p + (size_t)i * 3 * LCS_N_FOLD, pw.size() * 8, cudaMemcpyDeviceToHost));
        LCS_CUDA(ctx, cudaMemcpy(fq.data(), hb.frq.p + (size_t)i * 3 * LCS_N_FOLD, fq.size() * 4, cudaMemcpyDeviceToHost));
        LCS_CUDA(ctx, cudaMemcpy(spi.data(), hb.spi.p + (size_t)i * LCS_N_FOLD, spi.size() * 8, cudaMemcpyDeviceToHost));
        LCS_CUDA(ctx, cudaMemcpy(sg.data(), hb.single.p + (size_t)i * n_single, n_single * 4, cudaMemcpyDeviceToHost));
        calc_z_th1(spi.data(), LCS_N_FOLD, (uint16_t)g.n_comb_xc, (uint8_t)g.ds_comb_arm, z.data());
        auto at = [&](int t, int f, int idx) { return sg[((size_t)t * g.n_f_stride + f) * LCS_N_FOLD + idx]; };
        peak_search(pw.data(), fq.data(), z.data(), cfg.f.data(), cfg.fc_req, cfg.fc_prog, at, (uint8_t)g.ds_comb_arm, pk);
      } else {
        for (int k = 0; k < h_np[s][i]; k++) {
          const DevPeak& d = h_peaks[s][(size_t)i * SEARCH_MAX_PEAKS + k];
          lcs_cell c;
          lcs_cell_init(&c);
          c.fc_requested = cfg.fc_req;
          c.fc_programmed = cfg.fc_prog;
          c.pss_pow = d.pss_pow;
          c.ind = d.ind;
          c.freq = cfg.f[d.fi];
          c.n_id_2 = (int8_t)d.row;
          pk.push_back(c);
        }
      }
      ctx->chain_stream = ctx->streams[s];        // this chunk's stream is idle now: the per-peak kernels do not queue behind later chunks
      lcs_status rc = per_buffer(b0 + i, (const void*)(hb.iq.p + (size_t)i * g.n_cap * samp_bytes), pk);
      ctx->chain_stream = nullptr;
      if (rc != LCS_OK) return rc;
    }
    return LCS_OK;
  };
  // chunk k runs on stream k % NS; after issuing chunk k the host finishes chunk k - (NS - 1), so NS - 1 chunks are in
  // flight on the GPU while the oldest one's peaks are examined, and a stream's buffers are free again when it is reused
  const uint32_t n_chunks = (batch + chunk - 1) / chunk;
  for (uint32_t k = 0; k < n_chunks + (NS - 1); k++) {
    if (k < n_chunks) {
      lcs_status rc = issue(k * chunk, (int)(k % NS));
      if (rc != LCS_OK) return rc;
    }
    if (k >= (uint32_t)(NS - 1)) {
      const uint32_t kf = k - (NS - 1);
      lcs_status rc = finish(kf * chunk, (int)(kf % NS));
      if (rc != LCS_OK) return rc;
    }
  }
  return LCS_OK;
}

}  // namespace lcs

using namespace lcs;

// Multi-channel searcher: one plan per channel (planset.cu builds them in one launch), all channels of a chunk in one
// correlator launch.
struct lcs_sweep {
  lcs_ctx* ctx = nullptr;
  uint32_t n_cap = 0;
  PlanSet ps;
  lcs_xcorr_plan::HostBatchBufs hb[lcs_ctx::N_STREAMS];
  DevBuf<uint32_t> d_ident;              // 0, 1, 2, ...: plan of buffer b is b
  std::vector<uint32_t> h_ident;
};

static lcs_status sweep_prepare(lcs_sweep* sw, const std::vector<PlanCfg>& cfgs, uint8_t arm, bool want_fp32) {
  lcs_ctx* ctx = sw->ctx;
  cudaStream_t st = ctx->streams[0];
  // the previous call's kernels on the other streams may still read the old plans
  for (int i = 1; i < lcs_ctx::N_STREAMS; i++) LCS_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  lcs_status rc = planset_build(ctx, sw->ps, sw->n_cap, arm, cfgs, want_fp32, st);
  if (rc != LCS_OK) return rc;
  if (!want_fp32 && planset_resolve_kernel(sw->ps, LCS_KERNEL_AUTO, LCS_IQ_CU8) != LCS_KERNEL_TC) {
    rc = planset_build(ctx, sw->ps, sw->n_cap, arm, cfgs, true, st);       // this grid runs on the FP32 correlator
    if (rc != LCS_OK) return rc;
  }
  const uint32_t n = (uint32_t)cfgs.size();
  if (sw->h_ident.size() < n) {
    sw->h_ident.resize(n);
    for (uint32_t i = 0; i < n; i++) sw->h_ident[i] = i;
    LCS_CUDA(ctx, sw->d_ident.ensure(n));
    LCS_CUDA(ctx, cudaMemcpyAsync(sw->d_ident.p, sw->h_ident.data(), n * 4, cudaMemcpyHostToDevice, st));
  }
  // the plans are used from both streams
  int flag = 0;
  LCS_CUDA(ctx, cudaMemcpyAsync(&flag, sw->ps.d_flag.p, 4, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaStreamSynchronize(st));
  if (flag) { sw->ps.tc_ready = false; sw->ps.tc_why = "template digits outside the exact range of the integer formulation"; }
  return LCS_OK;
}

extern "C" {

lcs_status lcs_xcorr_peaks_batch_host(lcs_xcorr_plan* p, const void* iq_host, int iq_format, uint32_t batch, lcs_cell* peaks,
                                      uint32_t max_peaks, uint32_t* n_peaks) {
  if (!p) return fail(nullptr, LCS_ERR_ARG, "xcorr_peaks_batch_host: null plan");
  if (!iq_host || !n_peaks || (!peaks && max_peaks)) return fail(p->ctx, LCS_ERR_ARG, "xcorr_peaks_batch_host: null pointer");
  return search_chunks(p->ctx, p->ps, p->kernel, p->hb, iq_host, iq_format, batch, std::min<uint32_t>(p->max_batch, SEARCH_CHUNK), nullptr,
                       nullptr, [&](uint32_t b, const void*, const std::vector<lcs_cell>& pk) {
                         n_peaks[b] = (uint32_t)pk.size();
                         for (size_t k = 0; k < pk.size() && k < max_peaks; k++) peaks[(size_t)b * max_peaks + k] = pk[k];
                         return LCS_OK;
                       });
}

lcs_status lcs_cell_search_batch_cu8(lcs_xcorr_plan* p, const uint8_t* iq_host, uint32_t batch, lcs_cell* cells, uint32_t max_cells,
                                     uint32_t* n_cells) {
  if (!p) return fail(nullptr, LCS_ERR_ARG, "cell_search_batch_cu8: null plan");
  if (!iq_host || !n_cells || (!cells && max_cells)) return fail(p->ctx, LCS_ERR_ARG, "cell_search_batch_cu8: null pointer");
  lcs_ctx* ctx = p->ctx;
  const PlanCfg& cfg = p->ps.cfg[0];
  return search_chunks(ctx, p->ps, p->kernel, p->hb, iq_host, LCS_IQ_CU8, batch, std::min<uint32_t>(p->max_batch, SEARCH_CHUNK), nullptr, nullptr,
                       [&](uint32_t b, const void* d_cap, const std::vector<lcs_cell>& pk) {
                         uint32_t found = 0;
                         lcs_status rc = cell_chain_dev(ctx, d_cap, LCS_IQ_CU8, p->ps.geom.n_cap, pk, cfg.fc_req, cfg.fc_prog, cfg.fs_prog,
                                                        cells ? cells + (size_t)b * max_cells : nullptr, max_cells, &found);
                         n_cells[b] = found;
                         return rc;
                       });
}

// ---- multi-channel searcher -------------------------------------------------------------------------------------------
lcs_status lcs_sweep_create(lcs_ctx* ctx, uint32_t n_cap, lcs_sweep** out) {
  if (!ctx || !out) return fail(ctx, LCS_ERR_ARG, "sweep_create: null argument");
  if (n_cap < 136 + 100 + LCS_N_FOLD || n_cap < 273 + LCS_N_FOLD)
    return fail(ctx, LCS_ERR_ARG, "sweep_create: capture buffer shorter than one 5 ms half frame + margins");
  lcs_sweep* sw = new lcs_sweep();
  sw->ctx = ctx;
  sw->n_cap = n_cap;
  *out = sw;
  return LCS_OK;
}

void lcs_sweep_destroy(lcs_sweep* sw) {
  if (!sw) return;
  cudaSetDevice(sw->ctx->device);
  cudaDeviceSynchronize();
  delete sw;
}

// The per-centre-frequency loop of CellSearch (src/CellSearch.cpp:465-558) for n_ch capture buffers at once.
lcs_status lcs_sweep_search_cu8(lcs_sweep* sw, const uint8_t* iq_host, uint32_t n_ch, const double* fc_requested,
                                const double* fc_programmed, double fs_programmed, const double* f_search_set, uint32_t n_f,
                                lcs_cell* cells, uint32_t max_cells, uint32_t* n_cells) {
  if (!sw) return fail(nullptr, LCS_ERR_ARG, "sweep_search_cu8: null handle");
  lcs_ctx* ctx = sw->ctx;
  if (!iq_host || !fc_requested || !f_search_set || !n_cells || (!cells && max_cells)) return fail(ctx, LCS_ERR_ARG, "sweep_search_cu8: null pointer");
  if (n_ch == 0) return LCS_OK;
  std::vector<PlanCfg> cfgs(n_ch);
  for (uint32_t c = 0; c < n_ch; c++) {
    cfgs[c].fc_req = fc_requested[c];
    cfgs[c].fc_prog = fc_programmed ? fc_programmed[c] : fc_requested[c];
    cfgs[c].fs_prog = fs_programmed;
    cfgs[c].f.assign(f_search_set, f_search_set + n_f);
  }
  const uint8_t DS_COMB_ARM = 2;        // CellSearch.cpp:484
  lcs_status rc = sweep_prepare(sw, cfgs, DS_COMB_ARM, false);
  if (rc != LCS_OK) return rc;
  // chunks of 64 channels: 128 units of 38+ tiles keep every persistent correlator CTA busy for >30 tiles
  return search_chunks(ctx, sw->ps, LCS_KERNEL_AUTO, sw->hb, iq_host, LCS_IQ_CU8, n_ch, 64, sw->d_ident.p, sw->h_ident.data(),
                       [&](uint32_t b, const void* d_cap, const std::vector<lcs_cell>& pk) {
                         uint32_t found = 0;
                         const PlanCfg& cfg = sw->ps.cfg[b];
                         lcs_status r2 = cell_chain_dev(ctx, d_cap, LCS_IQ_CU8, sw->n_cap, pk, cfg.fc_req, cfg.fc_prog, cfg.fs_prog,
                                                        cells ? cells + (size_t)b * max_cells : nullptr, max_cells, &found);
                         n_cells[b] = found;
                         return r2;
                       });
}

// One searcher cycle (src/searcher_thread.cpp:95-232) for n_ch tracked channels at once: every channel is searched at its
// own single frequency offset.
lcs_status lcs_sweep_track_cu8(lcs_sweep* sw, const uint8_t* iq_host, uint32_t n_ch, const double* frequency_offset,
                               const double* fc_requested, const double* fc_programmed, double fs_programmed, const double* late,
                               const int32_t* tracked_n_id_cell, const uint32_t* n_tracked, uint32_t tracked_stride, lcs_cell* cells,
                               double* frame_timing, uint32_t max_cells, uint32_t* n_cells) {
  if (!sw) return fail(nullptr, LCS_ERR_ARG, "sweep_track_cu8: null handle");
  lcs_ctx* ctx = sw->ctx;
  if (!iq_host || !frequency_offset || !fc_requested || !n_cells || (max_cells && (!cells || !frame_timing)) ||
      (n_tracked && !tracked_n_id_cell))
    return fail(ctx, LCS_ERR_ARG, "sweep_track_cu8: null pointer");
  if (n_ch == 0) return LCS_OK;
  std::vector<PlanCfg> cfgs(n_ch);
  for (uint32_t c = 0; c < n_ch; c++) {
    cfgs[c].fc_req = fc_requested[c];
    cfgs[c].fc_prog = fc_programmed ? fc_programmed[c] : fc_requested[c];
    cfgs[c].fs_prog = fs_programmed;
    cfgs[c].f.assign(1, frequency_offset[c]);                                          // searcher_thread.cpp:96-98
  }
  lcs_status rc = sweep_prepare(sw, cfgs, 2, true);
  if (rc != LCS_OK) return rc;
  return search_chunks(ctx, sw->ps, LCS_KERNEL_AUTO, sw->hb, iq_host, LCS_IQ_CU8, n_ch, 64, sw->d_ident.p, sw->h_ident.data(),
                       [&](uint32_t b, const void* d_cap, const std::vector<lcs_cell>& pk) {
                         uint32_t found = 0;
                         const PlanCfg& cfg = sw->ps.cfg[b];
                         lcs_cell* out = cells ? cells + (size_t)b * max_cells : nullptr;
                         lcs_status r2 = cell_chain_dev(ctx, d_cap, LCS_IQ_CU8, sw->n_cap, pk, cfg.fc_req, cfg.fc_prog, cfg.fs_prog, out, max_cells,
                                                        &found, n_tracked ? tracked_n_id_cell + (size_t)b * tracked_stride : nullptr,
                                                        n_tracked ? n_tracked[b] : 0, true);
                         n_cells[b] = found;
                         const double k_factor = (cfg.fc_req - cfg.f[0]) / cfg.fc_prog;
                         for (uint32_t i = 0; i < found && i < max_cells; i++)                 // searcher_thread.cpp:214
                           frame_timing[(size_t)b * max_cells + i] = out[i].frame_start * (30720000.0 / 16) / (cfg.fs_prog * k_factor) + (late ? late[b] : 0.0);
                         return r2;
                       });
}

}  // extern "C"
