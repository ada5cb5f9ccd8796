// chain_gpu.cu - companion kernels of the correlator (FP64): batched PSS/SSS symbol extraction
// (fshift + 2-sample rotate + 128-point FFT), SSS channel-estimate combining, SSS maximum-likelihood
// detection over the 168 x {h1h2,h2h1} x {normal,extended} hypotheses, and OFDM time/frequency-grid
// extraction (whole-buffer FOC fused into 854 FFT-128).  Reference: src/searcher.cpp:516-935.
#include <cmath>

#include "chain_gpu.hpp"

namespace lcs {

static const double kPi = 3.14159265358979323846;
static const double kFsLte = 30720000.0;

// ---- sample loads as complex<double> ----
template <int FMT>
__device__ __forceinline__ double2 load_c(const void* __restrict__ base, size_t i);
template <>
__device__ __forceinline__ double2 load_c<LCS_IQ_C128>(const void* __restrict__ base, size_t i) {
  return __ldg(reinterpret_cast<const double2*>(base) + i);
}
template <>
__device__ __forceinline__ double2 load_c<LCS_IQ_CF32>(const void* __restrict__ base, size_t i) {
  float2 v = __ldg(reinterpret_cast<const float2*>(base) + i);
  return make_double2((double)v.x, (double)v.y);
}
template <>
__device__ __forceinline__ double2 load_c<LCS_IQ_CU8>(const void* __restrict__ base, size_t i) {
  uchar2 v = __ldg(reinterpret_cast<const uchar2*>(base) + i);
  return make_double2(((int)v.x - 127) / 128.0, ((int)v.y - 127) / 128.0);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 128-point forward FFT (e^{-j}), radix-2 DIT, 64 threads, data in shared memory in natural order
// on return.  `buf` must hold the input in BIT-REVERSED order on entry.
__device__ __forceinline__ void fft128_inplace(double2* buf, const double2* tw /*[64] e^{-j2pi k/128}*/, int tid) {
#pragma unroll
  for (int len = 2; len <= 128; len <<= 1) {
    const int half = len >> 1;
    const int k = tid & (half - 1);
    const int i = ((tid / half) * len) + k;
    const double2 w = tw[k * (128 / len)];
    const double2 u = buf[i], v = cmul(buf[i + half], w);
    buf[i] = make_double2(u.x + v.x, u.y + v.y);
    buf[i + half] = make_double2(u.x - v.x, u.y - v.y);
    __syncthreads();
  }
}
__device__ __forceinline__ int bitrev7(int i) { return (int)(__brev((unsigned)i) >> 25); }
__device__ __forceinline__ void make_twiddles(double2* tw, int tid) {
  double s, c;
  sincospi(-(double)tid / 64.0, &s, &c);  // e^{-j 2 pi tid/128}
  tw[tid] = make_double2(c, s);
}

// ---------------------------------------------------------------------------------------------
// extract_psss (searcher.cpp:516-530), batched: one block per 128-sample segment.
//   out[seg][62] = bins [-31..-1, 1..31] of dft( rotate_left_2( x[start..start+127] * e^{j k n} ) )
// with k = pi*foc_freq/(fs/2), n = 0..127 local (the reference shifts each segment from phase 0).
// ---------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(64) psss_kernel(const void* __restrict__ cap, const int* __restrict__ start,
                                                  const double k, double2* __restrict__ out) {
  __shared__ double2 buf[128];
  __shared__ double2 tw[64];
  const int tid = threadIdx.x, seg = blockIdx.x;
  make_twiddles(tw, tid);
  const size_t s0 = (size_t)start[seg];
  for (int n = tid; n < 128; n += 64) {
    double sn, cs;
    sincos(k * (double)n, &sn, &cs);
    const double2 x = cmul(load_c<FMT>(cap, s0 + n), make_double2(cs, sn));
    const int dst = (n + 126) & 127;  // rotate left by 2: b[i] = a[i+2]
    buf[bitrev7(dst)] = x;
  }
  __syncthreads();
  fft128_inplace(buf, tw, tid);
  const double sc = 1.0 / sqrt(128.0);
  if (tid < 62) {
    const int bin = tid < 31 ? 97 + tid : 1 + (tid - 31);
    out[(size_t)seg * 62 + tid] = make_double2(buf[bin].x * sc, buf[bin].y * sc);
  }
}

// ---------------------------------------------------------------------------------------------
// sss_detect_getce_sss (searcher.cpp:577-631) after the FFTs.  psss: [n_pss][3][62] =
// {PSS symbol, SSS symbol assuming extended CP, SSS symbol assuming normal CP}.  One block, thread t
// owns subcarrier t.  est: [h1_np 62][h2_np 62] doubles then c128 [h1_nrm][h2_nrm][h1_ext][h2_ext].
// ---------------------------------------------------------------------------------------------
constexpr int MAX_PSS = 64;
__global__ void __launch_bounds__(64) sss_getce_kernel(const double2* __restrict__ psss, const int n_pss,
                                                       const double2* __restrict__ pss_fd, double* __restrict__ est) {
  extern __shared__ double2 sm[];
  double2* h_raw = sm;                 // [n_pss][62]
  double2* h_sm = sm + n_pss * 62;     // [n_pss][62]
  __shared__ double np[MAX_PSS];
  const int t = threadIdx.x;
  if (t < 62) {
    const double2 pc = make_double2(pss_fd[t].x, -pss_fd[t].y);
    for (int k = 0; k < n_pss; k++) h_raw[k * 62 + t] = cmul(psss[((size_t)k * 3 + 0) * 62 + t], pc);
  }
  __syncthreads();
  if (t < 62) {
    const int lt = max(0, t - 6), rt = min(61, t + 6);
    for (int k = 0; k < n_pss; k++) {
      double2 s = make_double2(0, 0);
      for (int i = lt; i <= rt; i++) { s.x += h_raw[k * 62 + i].x; s.y += h_raw[k * 62 + i].y; }
      const double n = (double)(rt - lt + 1);
      h_sm[k * 62 + t] = make_double2(s.x / n, s.y / n);
    }
  }
  __syncthreads();
  if (t < n_pss) {  // noise power of PSS k (sigpower, dsp.h:23-29)
    double r = 0;
    for (int i = 0; i < 62; i++) {
      const double dx = h_sm[t * 62 + i].x - h_raw[t * 62 + i].x, dy = h_sm[t * 62 + i].y - h_raw[t * 62 + i].y;
      r += dx * dx + dy * dy;
    }
    np[t] = r / 62;
  }
  __syncthreads();
  if (t < 62) {
    double2* estc = reinterpret_cast<double2*>(est + 124);
    for (int half = 0; half < 2; half++) {
      double den = 0;
      double2 nrm = make_double2(0, 0), ext = make_double2(0, 0);
      for (int k = half; k < n_pss; k += 2) {
        const double inv = 1.0 / np[k];
        const double2 h = h_sm[k * 62 + t];
        den += (h.x * h.x + h.y * h.y) * inv;
        const double2 hc = make_double2(h.x * inv, -h.y * inv);
        const double2 a = cmul(hc, psss[((size_t)k * 3 + 2) * 62 + t]);
        const double2 b = cmul(hc, psss[((size_t)k * 3 + 1) * 62 + t]);
        nrm.x += a.x; nrm.y += a.y;
        ext.x += b.x; ext.y += b.y;
      }
      const double npe = 1.0 / (1.0 + den);
      est[half * 62 + t] = npe;
      estc[(0 + half) * 62 + t] = make_double2(npe * nrm.x, npe * nrm.y);
      estc[(2 + half) * 62 + t] = make_double2(npe * ext.x, npe * ext.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// sss_detect_ml (searcher.cpp:636-693): block = n_id_1, warp = hypothesis
// {nrm h1h2, nrm h2h1, ext h1h2, ext h2h1}.  sss_tab: int8 [168][3][2][62].
// ll: [4][168] = {nrm col0, nrm col1, ext col0, ext col1}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(128) sss_ml_kernel(const double* __restrict__ est, const signed char* __restrict__ sss_tab,
                                                     const int n_id_2, double* __restrict__ ll) {
  const int n1 = blockIdx.x, hyp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool swap = hyp & 1, is_ext = hyp >> 1;
  const double2* estc = reinterpret_cast<const double2*>(est + 124) + (is_ext ? 2 * 62 : 0);  // [h1][h2]
  const signed char* s0 = sss_tab + (((size_t)n1 * 3 + n_id_2) * 2 + 0) * 62;
  const signed char* s10 = s0 + 62;
  double ax = 0, ay = 0;
  for (int i = lane; i < 124; i += 32) {
    const int half = i >= 62, j = i - 62 * half;
    const double tr = (double)((half ^ swap) ? s10[j] : s0[j]);
    const double2 e = estc[i];
    ax += e.x * tr;   // conj(est)*try
    ay += -e.y * tr;
  }
  ax = warp_sum(ax);
  ay = warp_sum(ay);
  const double ang = atan2(ay, ax);
  double sn, cs;
  sincos(-ang, &sn, &cs);
  double sre = 0, sim = 0;
  for (int i = lane; i < 124; i += 32) {
    const int half = i >= 62, j = i - 62 * half;
    const double tr = (double)((half ^ swap) ? s10[j] : s0[j]);
    const double2 e = estc[i];
    const double dx = tr * cs - e.x, dy = tr * sn - e.y;
    const double npv = est[i];  // [h1_np][h2_np]
    sre += dx * dx / npv;
    sim += dy * dy / npv;
  }
  sre = warp_sum(sre);
  sim = warp_sum(sim);
  if (lane == 0) ll[hyp * 168 + n1] = -sre - sim;
}

// ---------------------------------------------------------------------------------------------
// extract_tfg (searcher.cpp:892-931): FOC of the whole buffer fused into the per-symbol FFT.
//   tfg[t][72] = bins [-36..-1,1..36] of dft( x[pos_t + n] * e^{j k (pos_t+n)} ) * e^{-j 2 pi late_t cn/128}
// ---------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(64) tfg_kernel(const void* __restrict__ cap, const int* __restrict__ pos,
                                                 const double* __restrict__ late, const double k,
                                                 double2* __restrict__ tfg) {
  __shared__ double2 buf[128];
  __shared__ double2 tw[64];
  const int tid = threadIdx.x, sym = blockIdx.x;
  make_twiddles(tw, tid);
  const size_t p0 = (size_t)pos[sym];
  for (int n = tid; n < 128; n += 64) {
    double sn, cs;
    sincos(k * (double)(p0 + n), &sn, &cs);
    buf[bitrev7(n)] = cmul(load_c<FMT>(cap, p0 + n), make_double2(cs, sn));
  }
  __syncthreads();
  fft128_inplace(buf, tw, tid);
  const double sc = 1.0 / sqrt(128.0);
  const double lt = late[sym];
  for (int i = tid; i < 72; i += 64) {
    const int bin = i < 36 ? 92 + i : 1 + (i - 36);
    const int cn = i < 36 ? i - 36 : i - 35;
    double sn, cs;
    sincos((-2.0 * kPi * lt / 128.0) * (double)cn, &sn, &cs);
    const double2 v = make_double2(buf[bin].x * sc, buf[bin].y * sc);
    tfg[(size_t)sym * 72 + i] = cmul(v, make_double2(cs, sn));
  }
}

#define DISPATCH(fmt, CALL)                               \
  do {                                                    \
    if ((fmt) == LCS_IQ_CU8) { CALL(LCS_IQ_CU8); }        \
    else if ((fmt) == LCS_IQ_CF32) { CALL(LCS_IQ_CF32); } \
    else { CALL(LCS_IQ_C128); }                           \
  } while (0)

// =============================================================================================
// Host drivers (device-resident capture buffer)
// =============================================================================================
static const signed char* sss_table_dev(lcs_ctx* ctx, ChainScratch& cs) {
  if (cs.d_sss_tab.p) return cs.d_sss_tab.p;
  std::vector<signed char> tab((size_t)168 * 3 * 2 * 62);
  int v[62];
  for (int n1 = 0; n1 < 168; n1++)
    for (int n2 = 0; n2 < 3; n2++)
      for (int s = 0; s < 2; s++) {
        sss_fd(n1, n2, s * 10, v);
        for (int i = 0; i < 62; i++) tab[(((size_t)n1 * 3 + n2) * 2 + s) * 62 + i] = (signed char)v[i];
      }
  if (cs.d_sss_tab.alloc(tab.size()) != cudaSuccess) return nullptr;
  cudaMemcpy(cs.d_sss_tab.p, tab.data(), tab.size(), cudaMemcpyHostToDevice);
  cd fd[62];
  cs.d_pss_fd.alloc(3 * 62);
  for (int t = 0; t < 3; t++) {
    pss_fd(t, fd);
    cudaMemcpy(cs.d_pss_fd.p + t * 62, fd, 62 * 16, cudaMemcpyHostToDevice);
  }
  (void)ctx;
  return cs.d_sss_tab.p;
}

static std::vector<double> mrange(double first, double incr, double last) {  // itpp_ext.cpp:97-108
  std::vector<double> r;
  auto sg = [](double x) { return (x > 0) - (x < 0); };
  if (sg(last - first) * sg(incr) >= 0) {
    const int n = (int)std::floor((last - first) / incr) + 1;
    for (int t = 0; t < n; t++) r.push_back(first + t * incr);
  }
  return r;
}
static inline double wrapd(double x, double sm, double lg) {  // macros.h:49 with itpp_ext.h:40-42
  const double n = lg - sm, k = x - sm;
  return (n == 0 ? k : k - n * (int)std::floor(k / n)) + sm;
}

static lcs_status run_psss(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, const std::vector<int>& starts,
                           double foc_freq, double fs_eff, cudaStream_t st) {
  LCS_CUDA(ctx, cs.d_starts.ensure(starts.size()));
  LCS_CUDA(ctx, cs.d_psss.ensure(starts.size() * 62));
  LCS_CUDA(ctx, cudaMemcpyAsync(cs.d_starts.p, starts.data(), starts.size() * 4, cudaMemcpyHostToDevice, st));
  const double k = kPi * foc_freq / (fs_eff / 2);  // dsp.h:42
#define CALL(F) psss_kernel<F><<<(unsigned)starts.size(), 64, 0, st>>>(d_cap, cs.d_starts.p, k, cs.d_psss.p)
  DISPATCH(fmt, CALL);
#undef CALL
  ctx->launches++;
  LCS_CUDA(ctx, cudaGetLastError());
  return LCS_OK;
}

lcs_status dev_sss_detect(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, uint32_t n_cap, const lcs_cell& cell,
                          double thresh2_n_sigma, double fc_req, double fc_prog, double fs_prog, lcs_cell& out,
                          SssDebugHost* dbg) {
  cudaStream_t st = ctx->streams[0];
  if (cell.n_id_2 < 0 || cell.n_id_2 > 2) return fail(ctx, LCS_ERR_ARG, "sss_detect: n_id_2 out of range");
  if (!sss_table_dev(ctx, cs)) return fail(ctx, LCS_ERR_CUDA, "sss table upload failed");
  // PSS positions with an SSS in front of them (searcher.cpp:549-563)
  double peak_loc = cell.ind;
  const double k_factor = (fc_req - cell.freq) / fc_prog;
  if (peak_loc + 9 < 162) peak_loc += 9600 * k_factor;
  const std::vector<double> locs = mrange(peak_loc, k_factor * 9600, (double)n_cap - 125 - 9);
  const int n_pss = (int)locs.size();
  if (n_pss < 1 || n_pss > MAX_PSS) return fail(ctx, LCS_ERR_RANGE, "sss_detect: unsupported number of PSS positions");
  std::vector<int> starts;
  for (int k = 0; k < n_pss; k++) {
    const long pss_dft = (long)std::rint(locs[k]) + 9 - 2;
    const long s[3] = {pss_dft, pss_dft - 128 - 32, pss_dft - 128 - 9};  // :579,:594,:596
    for (long v : s) {
      if (v < 0 || v + 128 > (long)n_cap) return fail(ctx, LCS_ERR_RANGE, "sss_detect: DFT window outside the capture buffer");
      starts.push_back((int)v);
    }
  }
  lcs_status rc = run_psss(ctx, cs, d_cap, fmt, starts, -cell.freq, fs_prog * k_factor, st);
  if (rc != LCS_OK) return rc;
  LCS_CUDA(ctx, cs.d_est.ensure(124 + 4 * 124));
  LCS_CUDA(ctx, cs.d_ll.ensure(4 * 168));
  if (!cs.getce_attr_set) {
    LCS_CUDA(ctx, cudaFuncSetAttribute(sss_getce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MAX_PSS * 62 * (int)sizeof(double2)));
    cs.getce_attr_set = true;
  }
  sss_getce_kernel<<<1, 64, (size_t)2 * n_pss * 62 * sizeof(double2), st>>>(cs.d_psss.p, n_pss, cs.d_pss_fd.p + cell.n_id_2 * 62, cs.d_est.p);
  sss_ml_kernel<<<168, 128, 0, st>>>(cs.d_est.p, cs.d_sss_tab.p, cell.n_id_2, cs.d_ll.p);
  ctx->launches += 2;
  LCS_CUDA(ctx, cudaGetLastError());
  std::vector<double> ll(4 * 168), est(124 + 4 * 124);
  LCS_CUDA(ctx, cudaMemcpyAsync(ll.data(), cs.d_ll.p, ll.size() * 8, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaMemcpyAsync(est.data(), cs.d_est.p, est.size() * 8, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaStreamSynchronize(st));
  // decisions (searcher.cpp:719-758)
  const double* nrm0 = &ll[0], *nrm1 = &ll[168], *ext0 = &ll[336], *ext1 = &ll[504];
  auto mx = [](const double* v) { return *std::max_element(v, v + 168); };
  const bool normal = std::max(mx(nrm0), mx(nrm1)) > std::max(mx(ext0), mx(ext1));
  const double* c0 = normal ? nrm0 : ext0, *c1 = normal ? nrm1 : ext1;
  const double fs_ratio = 16 / kFsLte * fs_prog;
  double frame_start = cell.ind + (128 + 9 - 960 - 2) * fs_ratio * k_factor;  // :735
  const double* col;
  if (mx(c0) > mx(c1)) col = c0;
  else { col = c1; frame_start += 9600 * k_factor * fs_ratio * k_factor; }  // :741
  frame_start = wrapd(frame_start, -0.5, (2 * 9600.0 - 0.5) * fs_ratio * k_factor);  // :743
  const int n_id_1 = (int)(std::max_element(col, col + 168) - col);
  const double lik_final = col[n_id_1];
  double sum = 0, sq = 0;  // IT++ mean / variance (N-1) over all 672 likelihoods
  for (double v : ll) { sum += v; sq += v * v; }
  const double mean = sum / 672, var = (sq - sum * sum / 672) / 671;
  out = cell;
  if (lik_final >= mean + std::pow(var, 0.5) * thresh2_n_sigma) {
    out.n_id_1 = n_id_1;
    out.cp_type = normal ? 1 : 2;
    out.frame_start = frame_start;
  }
  if (dbg) {
    dbg->est = est;
    dbg->ll = ll;
  }
  return LCS_OK;
}

lcs_status dev_pss_sss_foe(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, uint32_t n_cap, const lcs_cell& cell,
                           double fc_req, double fc_prog, double fs_prog, lcs_cell& out) {
  cudaStream_t st = ctx->streams[0];
  if (cell.n_id_1 < 0 || cell.n_id_1 > 167 || cell.n_id_2 < 0 || cell.n_id_2 > 2)
    return fail(ctx, LCS_ERR_ARG, "pss_sss_foe: cell id not set");
  const double k_factor = (fc_req - cell.freq) / fc_prog;
  const double fs_ratio = 16 / kFsLte * fs_prog;
  int dist;
  double first;
  if (cell.cp_type == 1) {
    dist = (int)std::rint((128 + 9) * fs_ratio * k_factor);  // :780
    first = cell.frame_start + (960 - 128 - 9 - 128) * fs_ratio * k_factor;
  } else if (cell.cp_type == 2) {
    dist = (int)std::rint((128 + 32) * k_factor);  // :783
    first = cell.frame_start + (960 - 128 - 32 - 128) * fs_ratio * k_factor;
  } else {
    return fail(ctx, LCS_ERR_ARG, "pss_sss_foe: cp_type unknown (reference throws \"Error... check code...\")");
  }
  int sn;
  first = wrapd(first, -0.5, 9600 * 2 - 0.5);
  if (first - 9600 * k_factor > -0.5) { first -= 9600 * k_factor; sn = 10; } else sn = 0;
  const std::vector<double> locs = mrange(first, 9600 * fs_ratio * k_factor, (double)((long)n_cap - 127 - dist - 100));
  const int n_sss = (int)locs.size();
  if (n_sss < 1) { out = cell; out.freq_fine = cell.freq; return LCS_OK; }   // reference: M stays 0, arg(0) = 0 (searcher.cpp:806,848)
  std::vector<int> starts;
  for (int k = 0; k < n_sss; k++) {
    const long s = (long)std::rint(locs[k]);
    if (s < 0 || s + dist + 128 > (long)n_cap) return fail(ctx, LCS_ERR_RANGE, "pss_sss_foe: DFT window outside the capture buffer");
    starts.push_back((int)(s + dist));  // PSS
    starts.push_back((int)s);           // SSS
  }
  lcs_status rc = run_psss(ctx, cs, d_cap, fmt, starts, -cell.freq, fs_prog * k_factor, st);
  if (rc != LCS_OK) return rc;
  std::vector<cd> bins(starts.size() * 62);
  LCS_CUDA(ctx, cudaMemcpyAsync(bins.data(), cs.d_psss.p, bins.size() * 16, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaStreamSynchronize(st));
  cd pfd[62];
  pss_fd(cell.n_id_2, pfd);
  sn = (1 - (sn / 10)) * 10;  // :800
  const double pa = kPi * -cell.freq / (kFsLte / 16 / 2) * -(double)dist;  // :832
  const cd ph(std::cos(pa), std::sin(pa));
  cd M = 0;
  for (int k = 0; k < n_sss; k++) {
    sn = (1 - (sn / 10)) * 10;
    cd h_raw[62], h_sm[62];
    for (int t = 0; t < 62; t++) h_raw[t] = bins[(size_t)(2 * k) * 62 + t] * std::conj(pfd[t]);
    for (int t = 0; t < 62; t++) {
      const int lt = std::max(0, t - 6), rt = std::min(61, t + 6);
      cd s = 0;
      for (int i = lt; i <= rt; i++) s += h_raw[i];
      h_sm[t] = s / (double)(rt - lt + 1);
    }
    double np = 0;
    for (int t = 0; t < 62; t++) np += std::norm(h_sm[t] - h_raw[t]);
    np /= 62;
    int sfd[62];
    sss_fd(cell.n_id_1, cell.n_id_2, sn, sfd);
    cd s = 0;
    for (int t = 0; t < 62; t++) {
      const cd sss = bins[(size_t)(2 * k + 1) * 62 + t] * ph * (double)sfd[t];
      const double a2 = std::norm(h_sm[t]);
      s += std::conj(sss) * h_raw[t] * (a2 / (2 * a2 * np + np * np));  // :836-843
    }
    M += s;
  }
  out = cell;
  out.freq_fine = cell.freq + std::arg(M) / (2 * kPi) / (1 / (fs_prog * k_factor) * dist);  // :848
  return LCS_OK;
}

lcs_status dev_extract_tfg(lcs_ctx* ctx, ChainScratch& cs, const void* d_cap, int fmt, uint32_t n_cap, const lcs_cell& cell,
                           double fc_req, double fc_prog, double fs_prog, std::vector<cd>& tfg, std::vector<double>& ts) {
  cudaStream_t st = ctx->streams[0];
  const double k_factor = (fc_req - cell.freq_fine) / fc_prog;  // :875
  const double fs_ratio = 16 / kFsLte * fs_prog;
  int n_symb;
  double loc;
  if (cell.cp_type == 1) { n_symb = 7; loc = cell.frame_start + 10 * fs_ratio * k_factor; }
  else if (cell.cp_type == 2) { n_symb = 6; loc = cell.frame_start + 32 * fs_ratio * k_factor; }
  else return fail(ctx, LCS_ERR_ARG, "extract_tfg: cp_type unknown (reference throws \"Check code...\")");
  if (!(std::isfinite(loc) && std::isfinite(cell.freq_fine))) return fail(ctx, LCS_ERR_ARG, "extract_tfg: frame_start / freq_fine not set");
  if (loc - .01 * fs_prog * k_factor > -0.5) loc -= .01 * fs_prog * k_factor;  // :887-889
  const int n_ofdm = 6 * 10 * 2 * n_symb + 2 * n_symb;
  ts.resize(n_ofdm);
  std::vector<int> pos(n_ofdm);
  std::vector<double> late(n_ofdm);
  int sym_num = 0;
  for (int t = 0; t < n_ofdm; t++) {  // :903-920 (same running sum as the reference)
    const double r = std::rint(loc);
    if (r < 0 || r + 128 > (double)n_cap) return fail(ctx, LCS_ERR_RANGE, "extract_tfg: DFT window outside the capture buffer");
    pos[t] = (int)r;
    ts[t] = loc;
    late[t] = r - loc;  // :925-928
    if (n_symb == 6) loc += (128 + 32) * fs_ratio * k_factor;
    else {
      loc += (sym_num == 6 ? (128 + 10) : (128 + 9)) * fs_ratio * k_factor;
      sym_num = (sym_num + 1) % 7;
    }
  }
  LCS_CUDA(ctx, cs.d_starts.ensure(n_ofdm));
  LCS_CUDA(ctx, cs.d_late.ensure(n_ofdm));
  LCS_CUDA(ctx, cs.d_tfg.ensure((size_t)n_ofdm * 72));
  LCS_CUDA(ctx, cudaMemcpyAsync(cs.d_starts.p, pos.data(), n_ofdm * 4, cudaMemcpyHostToDevice, st));
  LCS_CUDA(ctx, cudaMemcpyAsync(cs.d_late.p, late.data(), n_ofdm * 8, cudaMemcpyHostToDevice, st));
  const double k = kPi * -cell.freq_fine / ((fs_prog * k_factor) / 2);  // :892 via dsp.h:42
#define CALL(F) tfg_kernel<F><<<n_ofdm, 64, 0, st>>>(d_cap, cs.d_starts.p, cs.d_late.p, k, cs.d_tfg.p)
  DISPATCH(fmt, CALL);
#undef CALL
  ctx->launches++;
  LCS_CUDA(ctx, cudaGetLastError());
  tfg.resize((size_t)n_ofdm * 72);
  LCS_CUDA(ctx, cudaMemcpyAsync(tfg.data(), cs.d_tfg.p, tfg.size() * 16, cudaMemcpyDeviceToHost, st));
  LCS_CUDA(ctx, cudaStreamSynchronize(st));
  return LCS_OK;
}

}  // namespace lcs
