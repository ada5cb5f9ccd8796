// lcs_ctx.hpp - context / plan objects behind the opaque handles of include/lcs_b200.h.
#pragma once
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "lcs_internal.hpp"
#include "tc_layout.hpp"

namespace lcs {

template <typename T>
struct DevBuf {   // owning device allocation
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    n = count;
    return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  cudaError_t ensure(size_t count) { return (p && n >= count) ? cudaSuccess : alloc(count); }
};

template <class T>
struct PinBuf {   // owning page-locked host allocation (asynchronous copies need one)
  T* p = nullptr;
  size_t n = 0;
  PinBuf() {}
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  cudaError_t ensure(size_t count) {
    if (p && n >= count) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; }
    n = count;
    return cudaMallocHost((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
  }
};

// One search configuration: what xcorr_pss is called with besides the capture buffer (searcher.h:22-31).
struct PlanCfg {
  double fc_req = 0, fc_prog = 0, fs_prog = 0;
  std::vector<double> f;      // f_search_set
};

// A set of search configurations that share the capture-buffer shape: the operands of both correlator kernels for every
// plan, resident in HBM.  The integer geometry (fold offsets, pass tables) is computed on the host, the templates
// (conj(fshift(pss_td))/137 of searcher.cpp:145-151, their FP32 roundings and 24-bit digit planes) by one kernel
// (planset.cu) - a frequency sweep builds one plan per centre frequency in a single launch.
struct PlanSet {
  lcs_ctx* ctx = nullptr;
  XcorrGeom geom{};
  uint32_t n_plans = 0;
  std::vector<PlanCfg> cfg;
  std::vector<int> h_nf;
  // builder inputs
  DevBuf<double> d_cfg;            // [n_plans][4 + n_f_stride]: fc_req, fc_prog, fs_prog, n_f, f[]
  DevBuf<int> d_nf;                // [n_plans]
  // FP32 correlator operands
  bool has_fp32 = false;
  DevBuf<float4> d_w01;            // [n_plans][n_f_stride][140] (root0, root1)
  DevBuf<float2> d_w2;             // [n_plans][n_f_stride][140] root2
  DevBuf<int> d_soff;              // [n_plans][n_comb][n_f_stride]
  DevBuf<int> d_smin;              // [n_plans][n_comb][n_fchunk]
  // tensor-core correlator operands
  bool tc_ready = false;
  std::string tc_why;              // why not, when !tc_ready
  tc::Layout lay{16, 3, 2};
  uint32_t n_pass = 1;
  float inv_scale = 0;             // 1 / (S * 128)
  DevBuf<unsigned char> d_b;       // [n_plans][n_pass][lay.b_bytes()] int8 digit planes in UMMA core-matrix order
  DevBuf<float> d_corr;            // [n_plans][n_pass][2][npad]
  DevBuf<tc::PassGeo> d_geo;       // [n_plans][n_pass]
  DevBuf<int16_t> d_dsh;           // [n_plans][n_pass][M_MAX][npad] fold offset of the column minus the pass minimum
  DevBuf<int> d_flag;              // builder diagnostics (non-zero: a digit-plane bound was exceeded)
  PinBuf<unsigned char> h_stage;   // page-locked staging of the host-built tables
  cudaEvent_t staged = nullptr;    // the previous upload out of h_stage has completed
  ~PlanSet() { if (staged) cudaEventDestroy(staged); }
};

}  // namespace lcs

struct lcs_xcorr_plan;

struct lcs_ctx {
  int device = 0;
  int n_sm = 0;
  static constexpr int N_STREAMS = 3;           // chunks of the host-batch calls rotate over these
  cudaStream_t streams[N_STREAMS] = {nullptr, nullptr, nullptr};
  cudaStream_t chain_stream = nullptr;         // stream of the per-peak stages (NULL: streams[0]); the chunked search sets it to the idle stream of the chunk it examines
  std::string last_error;
  uint64_t launches = 0;
  std::vector<lcs_xcorr_plan*> cached_plans;   // for the plan-less drop-in calls
  // constants of the plan builder
  lcs::DevBuf<double> d_pss_td;                // [3][137] complex double (lte_lib.cpp:177-188)
  double tc_scale = 0;                         // power of two S: |template component| * S fits 24 bits for every offset
  // scratch of the drop-in host calls
  lcs::DevBuf<double> d_capbuf;                // c128 capture buffer (2 doubles / sample)
  lcs::DevBuf<float> d_single, d_ref, d_inc;
  lcs::DevBuf<double> d_pow, d_spi;
  lcs::DevBuf<int32_t> d_frq;
  lcs::DevBuf<double> d_work;                  // sss / tfg kernels
  lcs::DevBuf<unsigned char> d_cu8;
  lcs::DevBuf<int> d_flag8;                    // 8-bit exactness probe of lcs_xcorr_pss
  void* chain = nullptr;                       // lcs::ChainScratch (chain_api.cu), owned
};

struct lcs_xcorr_plan {
  lcs_ctx* ctx = nullptr;
  lcs::PlanSet ps;
  uint32_t max_batch = 1;
  int kernel = LCS_KERNEL_AUTO;
  // kernel timing hook
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool, ev_used;
  double ev_acc_ms = 0;                        // time of event pairs already harvested from ev_used
  uint64_t ev_acc_n = 0;
  // per-stream device buffers of the host-batch entry points
  struct HostBatchBufs {
    lcs::DevBuf<unsigned char> iq;
    lcs::DevBuf<float> single;
    lcs::DevBuf<double> pow, spi;
    lcs::DevBuf<int32_t> frq;
    // device peak search (search_batch.cu)
    lcs::DevBuf<double> work;
    lcs::DevBuf<unsigned char> peaks;
    lcs::DevBuf<int32_t> npeaks;
    lcs::PinBuf<unsigned char> h_peaks;   // page-locked landing zones of the peak lists
    lcs::PinBuf<int32_t> h_npeaks;
  } hb[lcs_ctx::N_STREAMS];
};

namespace lcs {

lcs_status fail(lcs_ctx* ctx, lcs_status st, const std::string& msg);
#define LCS_CUDA(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::lcs::fail((ctx), LCS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ---- planset.cu ----
// (Re)build the set for `cfgs` (all with the same n_cap / arm).  Asynchronous on `st` apart from the host-side geometry.
// want_fp32: also build the FP32 correlator's templates (skipped for 8-bit-only sweeps).
lcs_status planset_build(lcs_ctx* ctx, PlanSet& ps, uint32_t n_cap, uint8_t arm, const std::vector<PlanCfg>& cfgs,
                         bool want_fp32, cudaStream_t st);
// Which kernel AUTO resolves to for this set and input format.
int planset_resolve_kernel(const PlanSet& ps, int kernel, int iq_format);
// xcorr_pss for `batch` device-resident capture buffers: correlator + sp_est + delay spread / argmax.  d_buf_plan
// (device, [batch]) names the plan of every buffer (NULL: plan 0).
// ev: optional event pair recorded around the correlator kernel.
lcs_status planset_run(PlanSet& ps, int kernel, const void* d_iq, int iq_format, uint32_t batch, const uint32_t* d_buf_plan,
                       float* d_single, double* d_pow, int32_t* d_frq, double* d_spi, float* d_inc, cudaStream_t st,
                       const std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr);

lcs_status get_cached_plan(lcs_ctx* ctx, uint32_t n_cap, const double* f_search_set, uint32_t n_f, uint8_t arm,
                           double fc_req, double fc_prog, double fs_prog, lcs_xcorr_plan** out);
void chain_scratch_release(lcs_ctx* ctx);   // chain_api.cu
lcs_status cell_chain_dev(lcs_ctx* ctx, const void* d_cap, int fmt, uint32_t n_cap, const std::vector<lcs_cell>& pk, double fc_req,
                          double fc_prog, double fs_prog, lcs_cell* cells, uint32_t max_cells, uint32_t* n_cells,
                          const int32_t* tracked = nullptr, uint32_t n_tracked = 0, bool tracker_cycle = false);   // chain_api.cu
// ---- xcorr_tc.cu ----
lcs_status tc_init(lcs_ctx* ctx);           // one-time function attributes
int launch_xcorr_fold_tc(PlanSet& ps, const void* d_iq_cu8, uint32_t batch, const uint32_t* d_buf_plan,
                         float* d_single_planar, cudaStream_t st);
void tc_prof_dump();
// ---- lcs_api.cu ----
lcs_status plan_run_device(lcs_xcorr_plan* p, const void* d_iq, int iq_format, uint32_t batch, float* d_single, double* d_pow,
                           int32_t* d_frq, double* d_spi, float* d_inc, cudaStream_t st);

}  // namespace lcs
