// lcs_ctx.hpp - context / plan objects behind the opaque handles of include/lcs_b200.h.
#pragma once
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "lcs_internal.hpp"

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
struct PinBuf {   // owning page-locked host allocation (asynchronous device-to-host copies need one)
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

}  // namespace lcs

struct lcs_xcorr_plan;

struct lcs_ctx {
  int device = 0;
  int n_sm = 0;
  cudaStream_t streams[2] = {nullptr, nullptr};
  std::string last_error;
  uint64_t launches = 0;
  std::vector<lcs_xcorr_plan*> cached_plans;   // for the plan-less drop-in calls
  // scratch of the drop-in host calls
  lcs::DevBuf<double> d_capbuf;                // c128 capture buffer (2 doubles / sample)
  lcs::DevBuf<float> d_single, d_ref, d_inc;
  lcs::DevBuf<double> d_pow, d_spi;
  lcs::DevBuf<int32_t> d_frq;
  lcs::DevBuf<double> d_work;                  // sss / tfg kernels
  lcs::DevBuf<unsigned char> d_cu8;
};

struct lcs_xcorr_plan {
  lcs_ctx* ctx = nullptr;
  lcs::XcorrGeom geom{};
  std::vector<double> f_search_set;
  double fc_requested = 0, fc_programmed = 0, fs_programmed = 0;
  uint32_t max_batch = 1;
  int kernel = LCS_KERNEL_AUTO;
  std::vector<lcs::cd> h_w;       // [f][t][137] double-precision templates (host)
  std::vector<int> h_soff;        // [m][f]
  lcs::DevBuf<float4> d_w01;      // [n_f][140] (root0, root1)
  lcs::DevBuf<float2> d_w2;       // [n_f][140] root2
  lcs::DevBuf<int> d_soff, d_smin;
  lcs::DevBuf<double> d_sp_partial;
  // tensor-core path (xcorr_tc.cu)
  bool tc_ready = false;
  lcs::DevBuf<unsigned char> d_tc_a;   // packed template operand
  lcs::DevBuf<int> d_tc_meta;
  lcs::DevBuf<float> d_tc_scale;
  lcs::DevBuf<int16_t> d_tc_dsh;       // per-chunk fold-offset tables
  int tc_params[16] = {0};
  // kernel timing hook
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool, ev_used;
  // per-stream device buffers of lcs_xcorr_pss_batch_host
  struct HostBatchBufs {
    lcs::DevBuf<unsigned char> iq;
    lcs::DevBuf<float> single;
    lcs::DevBuf<double> pow, spi, sp_partial;
    lcs::DevBuf<int32_t> frq;
    // device peak search (search_batch.cu)
    lcs::DevBuf<double> work;
    lcs::DevBuf<unsigned char> peaks;
    lcs::DevBuf<int32_t> npeaks;
    lcs::PinBuf<unsigned char> h_peaks;   // page-locked landing zones of the peak lists
    lcs::PinBuf<int32_t> h_npeaks;
  } hb[2];
};

namespace lcs {

lcs_status fail(lcs_ctx* ctx, lcs_status st, const std::string& msg);
#define LCS_CUDA(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::lcs::fail((ctx), LCS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

lcs_status get_cached_plan(lcs_ctx* ctx, uint32_t n_cap, const double* f_search_set, uint32_t n_f, uint8_t arm,
                           double fc_req, double fc_prog, double fs_prog, lcs_xcorr_plan** out);
void chain_scratch_release(lcs_ctx* ctx);   // chain_api.cu
lcs_status cell_chain_dev(lcs_ctx* ctx, const void* d_cap, int fmt, uint32_t n_cap, const std::vector<lcs_cell>& pk, double fc_req,
                          double fc_prog, double fs_prog, lcs_cell* cells, uint32_t max_cells, uint32_t* n_cells,
                          const int32_t* tracked = nullptr, uint32_t n_tracked = 0);   // chain_api.cu
// xcorr_tc.cu
lcs_status tc_plan_setup(lcs_xcorr_plan* p);
void tc_prof_dump();
int launch_xcorr_fold_tc(lcs_xcorr_plan* p, const void* d_iq_cu8, uint32_t batch, float* d_single_planar, cudaStream_t st);

}  // namespace lcs
