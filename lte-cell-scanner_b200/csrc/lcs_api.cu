// lcs_api.cu - context, xcorr plan and the xcorr_pss entry points of the C ABI (include/lcs_b200.h).
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "lcs_ctx.hpp"

namespace lcs {

static std::string g_last_error;
static std::mutex g_err_mu;

lcs_status fail(lcs_ctx* ctx, lcs_status st, const std::string& msg) {
  {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_error = msg;
  }
  if (ctx) ctx->last_error = msg;
  return st;
}

// ---------------------------------------------------------------------------------------------
// Plan construction: templates, fold offsets, geometry.
// ---------------------------------------------------------------------------------------------
static lcs_status build_plan(lcs_ctx* ctx, uint32_t n_cap, const double* f_search_set, uint32_t n_f, uint8_t arm,
                             double fc_req, double fc_prog, double fs_prog, uint32_t max_batch, int kernel,
                             lcs_xcorr_plan** out) {
  if (!ctx || !f_search_set || !out) return fail(ctx, LCS_ERR_ARG, "xcorr plan: null argument");
  if (n_f == 0 || n_f > 4096) return fail(ctx, LCS_ERR_ARG, "xcorr plan: n_f out of range");
  if (n_cap < 136 + 100 + LCS_N_FOLD || n_cap < 273 + LCS_N_FOLD)
    return fail(ctx, LCS_ERR_ARG, "xcorr plan: capture buffer shorter than one 5 ms half frame + margins");
  if (arm > 64) return fail(ctx, LCS_ERR_ARG, "xcorr plan: ds_comb_arm out of range");
  if (max_batch == 0) max_batch = 1;
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));

  std::unique_ptr<lcs_xcorr_plan> p(new lcs_xcorr_plan());
  p->ctx = ctx;
  p->f_search_set.assign(f_search_set, f_search_set + n_f);
  p->fc_requested = fc_req;
  p->fc_programmed = fc_prog;
  p->fs_programmed = fs_prog;
  p->max_batch = max_batch;
  p->kernel = kernel;
  XcorrGeom& g = p->geom;
  g.n_cap = n_cap;
  g.n_f = n_f;
  g.ds_comb_arm = arm;
  const uint32_t n_lag = n_cap - 136;
  g.n_comb_xc = (n_lag - 100) / LCS_N_FOLD;      // searcher.cpp:276
  g.n_comb_sp = (n_cap - 136 - 137) / LCS_N_FOLD;  // searcher.cpp:194
  g.fw = n_f == 1 ? 1 : XC_FW;                    // searcher_thread.cpp:97-98 searches a single offset
  g.n_fchunk = (n_f + g.fw - 1) / g.fw;

  // Templates: conj(fshift(pss_td[t], f_off, fs_programmed*k_factor))/137  (searcher.cpp:145-151),
  // computed in double exactly like dsp.h:40-53 (cos/sin of k*t) and rounded once to fp32.
  cd td[3][137];
  for (int t = 0; t < 3; t++) pss_td(t, td[t]);
  const double kPi = 3.14159265358979323846;
  p->h_w.assign((size_t)n_f * 3 * 137, cd(0, 0));
  std::vector<float4> w01((size_t)n_f * XC_NTAP_PAD, make_float4(0, 0, 0, 0));
  std::vector<float2> w2((size_t)n_f * XC_NTAP_PAD, make_float2(0, 0));
  std::vector<int> soff((size_t)g.n_comb_xc * n_f);
  for (uint32_t f = 0; f < n_f; f++) {
    const double f_off = f_search_set[f];
    const double k_factor = (fc_req - f_off) / fc_prog;  // :147
    const double k = kPi * f_off / ((fs_prog * k_factor) / 2);
    for (int tap = 0; tap < 137; tap++) {
      const cd rot(std::cos(k * tap), std::sin(k * tap));
      cd w[3];
      for (int t = 0; t < 3; t++) {
        w[t] = std::conj(td[t][tap] * rot) / 137.0;
        p->h_w[((size_t)f * 3 + t) * 137 + tap] = w[t];
      }
      w01[(size_t)f * XC_NTAP_PAD + tap] = make_float4((float)w[0].real(), (float)w[0].imag(), (float)w[1].real(), (float)w[1].imag());
      w2[(size_t)f * XC_NTAP_PAD + tap] = make_float2((float)w[2].real(), (float)w[2].imag());
    }
    for (uint32_t m = 0; m < g.n_comb_xc; m++) {
      const double s = std::rint(m * .005 * k_factor * fs_prog);  // :298 (IT++ round_i == rint)
      if (s < 0 || s + (LCS_N_FOLD - 1) >= (double)n_lag)
        return fail(ctx, LCS_ERR_RANGE, "xcorr plan: fold offset runs past the correlation buffer (reference would read out of bounds)");
      soff[(size_t)m * n_f + f] = (int)s;
    }
  }
  p->h_soff = soff;
  std::vector<int> smin((size_t)g.n_comb_xc * g.n_fchunk);
  uint32_t max_spread = 0;
  for (uint32_t m = 0; m < g.n_comb_xc; m++)
    for (uint32_t c = 0; c < g.n_fchunk; c++) {
      int lo = INT32_MAX, hi = INT32_MIN;
      for (uint32_t f = c * g.fw; f < std::min(n_f, (c + 1) * g.fw); f++) {
        lo = std::min(lo, soff[(size_t)m * n_f + f]);
        hi = std::max(hi, soff[(size_t)m * n_f + f]);
      }
      smin[(size_t)m * g.n_fchunk + c] = lo;
      max_spread = std::max(max_spread, (uint32_t)(hi - lo));
    }
  g.max_spread = max_spread;
  g.tile_len = XC_TI * (XC_FW / g.fw) + XC_NTAP_PAD + max_spread + 8;
  const size_t smem = (size_t)g.fw * XC_NTAP_PAD * 24 + (size_t)g.tile_len * 8;
  if (smem > 100 * 1024)
    return fail(ctx, LCS_ERR_RANGE, "xcorr plan: frequency grid too sparse for one shared-memory tile (spread too large)");

  LCS_CUDA(ctx, p->d_w01.alloc(w01.size()));
  LCS_CUDA(ctx, p->d_w2.alloc(w2.size()));
  LCS_CUDA(ctx, p->d_soff.alloc(soff.size()));
  LCS_CUDA(ctx, p->d_smin.alloc(smin.size()));
  LCS_CUDA(ctx, cudaMemcpy(p->d_w01.p, w01.data(), w01.size() * sizeof(float4), cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, cudaMemcpy(p->d_w2.p, w2.data(), w2.size() * sizeof(float2), cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, cudaMemcpy(p->d_soff.p, soff.data(), soff.size() * sizeof(int), cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, cudaMemcpy(p->d_smin.p, smin.data(), smin.size() * sizeof(int), cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, p->d_sp_partial.alloc((size_t)max_batch * g.n_comb_sp * LCS_N_FOLD));
  lcs_status st = tc_plan_setup(p.get());
  if (st != LCS_OK) return st;
  *out = p.release();
  return LCS_OK;
}

static int resolve_kernel(const lcs_xcorr_plan* p, int iq_format) {
  if (p->kernel == LCS_KERNEL_FP32) return LCS_KERNEL_FP32;
  if (p->kernel == LCS_KERNEL_TC) return LCS_KERNEL_TC;
  // AUTO: the tensor-core kernel is exact only for 8-bit IQ; its cost is flat in n_f (one 128-row M tile per <=42
  // hypotheses) while the FP32 kernel's is proportional to n_f, so tiny grids (tracker mode, n_f=1) stay on FP32.
  return (iq_format == LCS_IQ_CU8 && p->tc_ready && p->geom.n_f >= 4) ? LCS_KERNEL_TC : LCS_KERNEL_FP32;
}

static lcs_status run_device(lcs_xcorr_plan* p, const void* d_iq, int iq_format, uint32_t batch, float* d_single,
                             double* d_pow, int32_t* d_frq, double* d_spi, float* d_inc, cudaStream_t st) {
  lcs_ctx* ctx = p->ctx;
  if (!d_iq || !d_single || !d_pow || !d_frq || !d_spi) return fail(ctx, LCS_ERR_ARG, "xcorr_pss_device: null pointer");
  if (batch == 0 || batch > p->max_batch) return fail(ctx, LCS_ERR_ARG, "xcorr_pss_device: batch exceeds plan max_batch");
  if (iq_format != LCS_IQ_CF32 && iq_format != LCS_IQ_CU8 && iq_format != LCS_IQ_C128)
    return fail(ctx, LCS_ERR_ARG, "xcorr_pss_device: bad iq_format");
  const int kern = resolve_kernel(p, iq_format);
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  if (p->timing) {
    if (p->ev_pool.empty()) {
      LCS_CUDA(ctx, cudaEventCreate(&ev.first));
      LCS_CUDA(ctx, cudaEventCreate(&ev.second));
    } else {
      ev = p->ev_pool.back();
      p->ev_pool.pop_back();
    }
    LCS_CUDA(ctx, cudaEventRecord(ev.first, st));
  }
  if (kern == LCS_KERNEL_TC) {
    if (iq_format != LCS_IQ_CU8) return fail(ctx, LCS_ERR_ARG, "tensor-core correlator needs LCS_IQ_CU8 input");
    if (!p->tc_ready) return fail(ctx, LCS_ERR_STATE, "tensor-core correlator not available for this plan");
    ctx->launches += launch_xcorr_fold_tc(p, d_iq, batch, d_single, st);
  } else {
    ctx->launches += launch_xcorr_fold_fp32(p->geom, d_iq, iq_format, batch, p->d_w01.p, p->d_w2.p, p->d_soff.p,
                                            p->d_smin.p, d_single, st);
  }
  if (p->timing) {
    LCS_CUDA(ctx, cudaEventRecord(ev.second, st));
    p->ev_used.push_back(ev);
  }
  ctx->launches += launch_sp_partial(p->geom, d_iq, iq_format, batch, p->d_sp_partial.p, st);
  ctx->launches += launch_epilogue(p->geom, batch, d_single, p->d_sp_partial.p, d_pow, d_frq, d_spi, d_inc, st);
  LCS_CUDA(ctx, cudaGetLastError());
  return LCS_OK;
}

}  // namespace lcs

using namespace lcs;

extern "C" {

const char* lcs_version(void) { return "lcs_b200 0.1 (sm_100a)"; }

lcs_status lcs_ctx_create(int device, lcs_ctx** out) {
  if (!out) return fail(nullptr, LCS_ERR_ARG, "ctx_create: null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, LCS_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, LCS_ERR_ARG, "ctx_create: device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, LCS_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, LCS_ERR_CUDA, "device is not compute capability 10.x (kernels are built for sm_100a only)");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, LCS_ERR_CUDA, cudaGetErrorString(e));
  lcs_ctx* c = new lcs_ctx();
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  for (int i = 0; i < 2; i++) {
    if (cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return fail(nullptr, LCS_ERR_CUDA, "stream creation failed");
    }
  }
  *out = c;
  return LCS_OK;
}

void lcs_ctx_destroy(lcs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto* p : ctx->cached_plans) delete p;
  ctx->cached_plans.clear();
  chain_scratch_release(ctx);
  for (int i = 0; i < 2; i++)
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
  delete ctx;
}

const char* lcs_last_error(const lcs_ctx* ctx) {
  if (ctx) return ctx->last_error.c_str();
  std::lock_guard<std::mutex> lk(g_err_mu);
  static thread_local std::string copy;
  copy = g_last_error;
  return copy.c_str();
}

uint64_t lcs_launch_count(const lcs_ctx* ctx) { return ctx ? ctx->launches : 0; }

void lcs_cell_init(lcs_cell* c) {  // Cell::Cell(), reference src/common.cpp:36-56
  if (!c) return;
  c->fc_requested = c->fc_programmed = c->pss_pow = NAN;
  c->ind = -1;
  c->freq = NAN;
  c->n_id_2 = -1;
  c->n_id_1 = -1;
  c->cp_type = 0;
  c->frame_start = c->freq_fine = c->freq_superfine = NAN;
  c->n_ports = c->n_rb_dl = -1;
  c->phich_duration = c->phich_resource = 0;
  c->sfn = -1;
}

lcs_status lcs_xcorr_plan_create(lcs_ctx* ctx, uint32_t n_cap, const double* f_search_set, uint32_t n_f,
                                 uint8_t ds_comb_arm, double fc_requested, double fc_programmed, double fs_programmed,
                                 uint32_t max_batch, int kernel, lcs_xcorr_plan** plan) {
  return build_plan(ctx, n_cap, f_search_set, n_f, ds_comb_arm, fc_requested, fc_programmed, fs_programmed, max_batch,
                    kernel, plan);
}

void lcs_xcorr_plan_destroy(lcs_xcorr_plan* plan) {
  if (!plan) return;
  tc_prof_dump();
  cudaSetDevice(plan->ctx->device);
  cudaDeviceSynchronize();
  for (auto& ev : plan->ev_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto& ev : plan->ev_used) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  delete plan;
}

lcs_status lcs_xcorr_plan_timing_enable(lcs_xcorr_plan* p, int enable) {
  if (!p) return fail(nullptr, LCS_ERR_ARG, "timing_enable: null plan");
  p->timing = enable != 0;
  return LCS_OK;
}
lcs_status lcs_xcorr_plan_timing_read(lcs_xcorr_plan* p, double* kernel_ms, uint64_t* launches) {
  if (!p || !kernel_ms || !launches) return fail(nullptr, LCS_ERR_ARG, "timing_read: null argument");
  double tot = 0;
  for (auto& ev : p->ev_used) {
    LCS_CUDA(p->ctx, cudaEventSynchronize(ev.second));
    float ms = 0;
    LCS_CUDA(p->ctx, cudaEventElapsedTime(&ms, ev.first, ev.second));
    tot += ms;
    p->ev_pool.push_back(ev);
  }
  *kernel_ms = tot;
  *launches = p->ev_used.size();
  p->ev_used.clear();
  return LCS_OK;
}

uint16_t lcs_xcorr_plan_n_comb_xc(const lcs_xcorr_plan* p) { return p ? (uint16_t)p->geom.n_comb_xc : 0; }
uint16_t lcs_xcorr_plan_n_comb_sp(const lcs_xcorr_plan* p) { return p ? (uint16_t)p->geom.n_comb_sp : 0; }
int lcs_xcorr_plan_kernel(const lcs_xcorr_plan* p, int iq_format) { return p ? resolve_kernel(p, iq_format) : 0; }

lcs_status lcs_xcorr_pss_device(lcs_xcorr_plan* plan, const void* d_iq, int iq_format, uint32_t batch,
                                float* d_single_planar, double* d_pow, int32_t* d_frq, double* d_sp_incoherent,
                                float* d_incoherent_planar, void* stream) {
  if (!plan) return fail(nullptr, LCS_ERR_ARG, "xcorr_pss_device: null plan");
  return run_device(plan, d_iq, iq_format, batch, d_single_planar, d_pow, d_frq, d_sp_incoherent, d_incoherent_planar,
                    (cudaStream_t)stream);
}

// Host-buffer batched call: chunks of the batch alternate between the context's two streams so
// that H2D(i+1) and D2H(i-1) overlap the kernels of chunk i.
lcs_status lcs_xcorr_pss_batch_host(lcs_xcorr_plan* p, const void* h_iq, int iq_format, uint32_t batch,
                                    float* h_single, double* h_pow, int32_t* h_frq, double* h_spi) {
  if (!p) return fail(nullptr, LCS_ERR_ARG, "xcorr_pss_batch_host: null plan");
  lcs_ctx* ctx = p->ctx;
  if (!h_iq || !h_pow || !h_frq || !h_spi) return fail(ctx, LCS_ERR_ARG, "xcorr_pss_batch_host: null pointer");
  if (batch == 0) return LCS_OK;
  const XcorrGeom& g = p->geom;
  const size_t samp_bytes = iq_format == LCS_IQ_CU8 ? 2 : (iq_format == LCS_IQ_CF32 ? 8 : (iq_format == LCS_IQ_C128 ? 16 : 0));
  if (!samp_bytes) return fail(ctx, LCS_ERR_ARG, "xcorr_pss_batch_host: bad iq_format");
  LCS_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t chunk = std::min<uint32_t>(std::min<uint32_t>(p->max_batch, 8u), batch);
  const size_t n_single = (size_t)3 * g.n_f * LCS_N_FOLD;
  for (int s = 0; s < 2; s++) {
    LCS_CUDA(ctx, p->hb[s].iq.ensure((size_t)chunk * g.n_cap * 16));
    LCS_CUDA(ctx, p->hb[s].single.ensure(chunk * n_single));
    LCS_CUDA(ctx, p->hb[s].pow.ensure((size_t)chunk * 3 * LCS_N_FOLD));
    LCS_CUDA(ctx, p->hb[s].frq.ensure((size_t)chunk * 3 * LCS_N_FOLD));
    LCS_CUDA(ctx, p->hb[s].spi.ensure((size_t)chunk * LCS_N_FOLD));
    LCS_CUDA(ctx, p->hb[s].sp_partial.ensure((size_t)chunk * g.n_comb_sp * LCS_N_FOLD));
  }
  int s = 0;
  for (uint32_t b0 = 0; b0 < batch; b0 += chunk, s ^= 1) {
    const uint32_t nb = std::min(chunk, batch - b0);
    cudaStream_t st = ctx->streams[s];
    auto& hb = p->hb[s];
    LCS_CUDA(ctx, cudaMemcpyAsync(hb.iq.p, (const char*)h_iq + (size_t)b0 * g.n_cap * samp_bytes,
                                  (size_t)nb * g.n_cap * samp_bytes, cudaMemcpyHostToDevice, st));
    // each stream needs its own sp_partial scratch
    double* saved = p->d_sp_partial.p;
    p->d_sp_partial.p = hb.sp_partial.p;
    lcs_status rc = run_device(p, hb.iq.p, iq_format, nb, hb.single.p, hb.pow.p, hb.frq.p, hb.spi.p, nullptr, st);
    p->d_sp_partial.p = saved;
    if (rc != LCS_OK) return rc;
    if (h_single)
      LCS_CUDA(ctx, cudaMemcpyAsync(h_single + (size_t)b0 * n_single, hb.single.p, (size_t)nb * n_single * 4, cudaMemcpyDeviceToHost, st));
    LCS_CUDA(ctx, cudaMemcpyAsync(h_pow + (size_t)b0 * 3 * LCS_N_FOLD, hb.pow.p, (size_t)nb * 3 * LCS_N_FOLD * 8, cudaMemcpyDeviceToHost, st));
    LCS_CUDA(ctx, cudaMemcpyAsync(h_frq + (size_t)b0 * 3 * LCS_N_FOLD, hb.frq.p, (size_t)nb * 3 * LCS_N_FOLD * 4, cudaMemcpyDeviceToHost, st));
    LCS_CUDA(ctx, cudaMemcpyAsync(h_spi + (size_t)b0 * LCS_N_FOLD, hb.spi.p, (size_t)nb * LCS_N_FOLD * 8, cudaMemcpyDeviceToHost, st));
  }
  LCS_CUDA(ctx, cudaStreamSynchronize(ctx->streams[0]));
  LCS_CUDA(ctx, cudaStreamSynchronize(ctx->streams[1]));
  return LCS_OK;
}

// Drop-in for searcher.h:22-41.
lcs_status lcs_xcorr_pss(lcs_ctx* ctx, const double* capbuf, uint32_t n_cap, const double* f_search_set, uint32_t n_f,
                         uint8_t ds_comb_arm, double fc_requested, double fc_programmed, double fs_programmed,
                         double* pow, int32_t* frq, float* single, float* incoherent, double* sp_incoherent, float* xc,
                         double* sp, uint16_t* n_comb_xc, uint16_t* n_comb_sp) {
  if (!ctx) return fail(nullptr, LCS_ERR_ARG, "xcorr_pss: null context");
  if (!capbuf || !f_search_set || !pow || !frq || !single || !sp_incoherent)
    return fail(ctx, LCS_ERR_ARG, "xcorr_pss: null pointer");
  lcs_xcorr_plan* p = nullptr;
  lcs_status rc = get_cached_plan(ctx, n_cap, f_search_set, n_f, ds_comb_arm, fc_requested, fc_programmed, fs_programmed, &p);
  if (rc != LCS_OK) return rc;
  const XcorrGeom& g = p->geom;
  cudaStream_t st = ctx->streams[0];
  const size_t n_single = (size_t)3 * n_f * LCS_N_FOLD;
  LCS_CUDA(ctx, ctx->d_capbuf.ensure((size_t)n_cap * 2));
  LCS_CUDA(ctx, ctx->d_single.ensure(n_single));
  LCS_CUDA(ctx, ctx->d_ref.ensure(n_single));
  LCS_CUDA(ctx, ctx->d_inc.ensure(n_single));
  LCS_CUDA(ctx, ctx->d_pow.ensure(3 * LCS_N_FOLD));
  LCS_CUDA(ctx, ctx->d_frq.ensure(3 * LCS_N_FOLD));
  LCS_CUDA(ctx, ctx->d_spi.ensure(LCS_N_FOLD));
  LCS_CUDA(ctx, cudaMemcpyAsync(ctx->d_capbuf.p, capbuf, (size_t)n_cap * 16, cudaMemcpyHostToDevice, st));
  rc = run_device(p, ctx->d_capbuf.p, LCS_IQ_C128, 1, ctx->d_single.p, ctx->d_pow.p, ctx->d_frq.p, ctx->d_spi.p,
                  incoherent ? ctx->d_inc.p : nullptr, st);
  if (rc != LCS_OK) return rc;
  // reference layouts: vf3d [t][idx][f]; mat(3,9600) column-major
  ctx->launches += launch_planar_to_ref(g, ctx->d_single.p, ctx->d_ref.p, st);
  LCS_CUDA(ctx, cudaMemcpyAsync(single, ctx->d_ref.p, n_single * 4, cudaMemcpyDeviceToHost, st));
  if (incoherent) {
    LCS_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += launch_planar_to_ref(g, ctx->d_inc.p, ctx->d_ref.p, st);
    LCS_CUDA(ctx, cudaMemcpyAsync(incoherent, ctx->d_ref.p, n_single * 4, cudaMemcpyDeviceToHost, st));
  }
  std::vector<double> hpow(3 * LCS_N_FOLD);
  std::vector<int32_t> hfrq(3 * LCS_N_FOLD);
  LCS_CUDA(ctx, cudaMemcpyAsync(hpow.data(), ctx->d_pow.p, hpow.size() * 8, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaMemcpyAsync(hfrq.data(), ctx->d_frq.p, hfrq.size() * 4, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaMemcpyAsync(sp_incoherent, ctx->d_spi.p, LCS_N_FOLD * 8, cudaMemcpyDeviceToHost, st));
  if (xc) {
    const size_t n_xc = (size_t)3 * (n_cap - 136) * n_f;
    DevBuf<float2> d_xc;
    LCS_CUDA(ctx, d_xc.alloc(n_xc));
    ctx->launches += launch_xc_debug(g, ctx->d_capbuf.p, LCS_IQ_C128, p->d_w01.p, p->d_w2.p, d_xc.p, st);
    LCS_CUDA(ctx, cudaMemcpyAsync(xc, d_xc.p, n_xc * 8, cudaMemcpyDeviceToHost, st));
    LCS_CUDA(ctx, cudaStreamSynchronize(st));
  }
  if (sp) {
    DevBuf<double> d_sp;
    LCS_CUDA(ctx, d_sp.alloc((size_t)g.n_comb_sp * LCS_N_FOLD));
    ctx->launches += launch_sp_debug(g, ctx->d_capbuf.p, LCS_IQ_C128, d_sp.p, st);
    LCS_CUDA(ctx, cudaMemcpyAsync(sp, d_sp.p, (size_t)g.n_comb_sp * LCS_N_FOLD * 8, cudaMemcpyDeviceToHost, st));
    LCS_CUDA(ctx, cudaStreamSynchronize(st));
  }
  LCS_CUDA(ctx, cudaStreamSynchronize(st));
  for (int t = 0; t < 3; t++)
    for (int k = 0; k < LCS_N_FOLD; k++) {
      pow[(size_t)k * 3 + t] = hpow[(size_t)t * LCS_N_FOLD + k];
      frq[(size_t)k * 3 + t] = hfrq[(size_t)t * LCS_N_FOLD + k];
    }
  if (n_comb_xc) *n_comb_xc = (uint16_t)g.n_comb_xc;
  if (n_comb_sp) *n_comb_sp = (uint16_t)g.n_comb_sp;
  return LCS_OK;
}

}  // extern "C"

namespace lcs {
lcs_status get_cached_plan(lcs_ctx* ctx, uint32_t n_cap, const double* f_search_set, uint32_t n_f, uint8_t arm,
                           double fc_req, double fc_prog, double fs_prog, lcs_xcorr_plan** out) {
  for (auto* q : ctx->cached_plans) {
    if (q->geom.n_cap == n_cap && q->geom.n_f == n_f && q->geom.ds_comb_arm == arm && q->fc_requested == fc_req &&
        q->fc_programmed == fc_prog && q->fs_programmed == fs_prog &&
        std::memcmp(q->f_search_set.data(), f_search_set, n_f * sizeof(double)) == 0) {
      *out = q;
      return LCS_OK;
    }
  }
  lcs_xcorr_plan* p = nullptr;
  static const bool trace = std::getenv("LCS_TRACE_PLANS") != nullptr;     // debug aid: host cost of plan turnover
  const auto t0 = std::chrono::steady_clock::now();
  lcs_status rc = build_plan(ctx, n_cap, f_search_set, n_f, arm, fc_req, fc_prog, fs_prog, 1, LCS_KERNEL_AUTO, &p);
  if (rc != LCS_OK) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  if (ctx->cached_plans.size() >= 8) {
    delete ctx->cached_plans.front();
    ctx->cached_plans.erase(ctx->cached_plans.begin());
  }
  if (trace)
    std::fprintf(stderr, "[lcs] plan build %.3f ms, evict %.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  ctx->cached_plans.push_back(p);
  *out = p;
  return LCS_OK;
}
}  // namespace lcs
