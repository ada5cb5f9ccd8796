// chain_gpu.hpp - drivers of the companion kernels (chain_gpu.cu) on a device-resident capture buffer.
#pragma once
#include "chain_host.hpp"
#include "lcs_ctx.hpp"

namespace lcs {

struct ChainScratch {
  DevBuf<signed char> d_sss_tab;   // [168][3][2][62] +-1
  bool getce_attr_set = false;
  DevBuf<double2> d_pss_fd;        // [3][62]
  DevBuf<int> d_starts;
  DevBuf<double2> d_psss;          // [n_seg][62]
  DevBuf<double> d_est;            // [124] np + 4x62 complex
  DevBuf<double> d_ll;             // [4][168]
  DevBuf<double> d_late;
  DevBuf<double2> d_tfg;           // [n_ofdm][72]
};
ChainScratch& chain_scratch(lcs_ctx* ctx);

struct SssDebugHost {
  std::vector<double> est;  // [h1_np 62][h2_np 62][h1_nrm 124][h2_nrm 124][h1_ext 124][h2_ext 124]
  std::vector<double> ll;   // [nrm col0 168][nrm col1][ext col0][ext col1]
};

lcs_status dev_sss_detect(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, uint32_t n_cap, const lcs_cell& cell,
                          double thresh2_n_sigma, double fc_req, double fc_prog, double fs_prog, lcs_cell& out,
                          SssDebugHost* dbg);
lcs_status dev_pss_sss_foe(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, uint32_t n_cap, const lcs_cell& cell,
                           double fc_req, double fc_prog, double fs_prog, lcs_cell& out);
lcs_status dev_extract_tfg(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, uint32_t n_cap, const lcs_cell& cell,
                           double fc_req, double fc_prog, double fs_prog, std::vector<cd>& tfg_rowmajor,
                           std::vector<double>& ts);

}  // namespace lcs
