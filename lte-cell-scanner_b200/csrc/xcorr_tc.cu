// xcorr_tc.cu - PSS correlator on the 5th-generation tensor cores (tcgen05 / TMEM), exact for 8-bit IQ.
//
// The sliding correlation is a Toeplitz GEMM: D[lag, template] = sum_j z[2*lag + j] * W[template, j],
// j = 0..273 over the interleaved I/Q byte stream z of the capture buffer (rtl-sdr wire format,
// reference src/capbuf.cpp:157-181).  Everything is done in EXACT integer arithmetic:
//
//   * IQ bytes v are used as signed x' = v-128 (a XOR with 0x80; the true sample is (x'+1)/128),
//   * each template component W (double, conj(fshift(pss_td))/137 of searcher.cpp:145-151) is scaled by
//     a power of two S and rounded to a 24-bit integer, split into three balanced base-256 digits
//     W*S = 65536 a0 + 256 a1 + a2, a_j in [-128,127] -> three int8 B operands,
//   * tcgen05.mma kind::i8 (s8 x s8 -> s32 accumulators in TMEM): |sum| <= 274*128*128 < 2^23, no overflow,
//   * real part uses the byte stream as is, the imaginary part a second stream with every (I,Q) pair
//     replaced by (Q, ~I)  (~I = -I'-1): sum a[2m]*Q' + a[2m+1]*(-I'-1) with the same template rows
//     a[2m] = Re W, a[2m+1] = -Im W.
//
// Operand roles: the 128 LAGS of a sub-tile are the M dimension (TMEM lanes), the <= 96 TEMPLATES of a chunk
// (3 PSS roots x <= 32 frequency hypotheses) the N dimension (TMEM columns).  With the lags on the lanes
//   * all four warp schedulers of the SM share the epilogue evenly (each lane quarter carries lags),
//   * a template's fold offset (its k_factor, searcher.cpp:298) is uniform across a warp: the read-modify-write
//     of the incoherent sum touches 32 consecutive floats - conflict-free for any offset,
//   * the per-template constants are warp-uniform operands.
//
// The Toeplitz (Hankel) A operand is never materialised per lag: an "expanded" tile P[u][r][16 B] = z[16u+2r ..+15]
// is built once per 256-lag tile in shared memory (8x expansion of ~0.8 KB); block u is exactly the
// 8-row x 16-byte K-major core matrix of (row group g, K chunk c) for every g+c = u, so one UMMA
// shared-memory descriptor with LBO = SBO = 128 B addresses the whole Hankel tile.  The template digit planes
// stay resident in shared memory in core-matrix order (LBO 128 B, SBO 2304 B).
//
// Per CTA (persistent, one per SM): warp 0 builds P tiles, warp 1 issues the MMAs (one elected lane; per sub-tile and
// re/im part a "wide" job - digit planes 0|1 side by side, N = 2*npad, 9 UTCIMMA - into one of two 192-column TMEM slots
// and a "narrow" job - digit plane 2, N = npad - into a third slot), the epilogue warps (4 per group of NC template
// columns, one per TMEM lane quarter) read the slots back (tcgen05.ld), recombine the digits, turn them into |xc|^2 and
// fold the 15 half frames into per-template accumulators in shared memory.
// Output: xc_incoherent_single, planar [batch][3][n_f][9600] float.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

#include "lcs_ctx.hpp"

// Build with -DLCS_TC_PROFILE=1 to collect per-CTA cycle counters of the pipeline stages (printed when a plan is
// destroyed with LCS_TC_PROF=1 in the environment); off by default so that the clock reads cost nothing.
#ifndef LCS_TC_PROFILE
#define LCS_TC_PROFILE 0
#endif
#ifndef LCS_TC_DBG
#define LCS_TC_DBG 0    // 1: honour TcParams::dbg (timing experiments that skip epilogue stages; wrong results)
#endif
#if LCS_TC_PROFILE
#define TC_CLK() clock64()
#else
#define TC_CLK() 0ll
#endif

namespace lcs {

namespace tc {
constexpr int NT = 256;            // lags per tile
constexpr int NSUBL = 128;         // lags per MMA (UMMA M)
constexpr int NSUB = NT / NSUBL;   // MMA sub-tiles per tile
constexpr int KB = 288;            // K in bytes: 274 interleaved I/Q taps padded to a multiple of 32
constexpr int KSTEPS = KB / 32;    // UTCIMMA K = 32 bytes
constexpr int KCHUNKS = KB / 16;   // 16-byte K chunks (core-matrix columns)
constexpr int NBLK = NT / 8 + KB / 16 - 1;   // 49 expanded blocks of 128 B per tile
constexpr int P_BYTES = NBLK * 128;          // one variant of one stage
constexpr int N_MAX = 96;          // template columns per chunk (UMMA N, a multiple of 32): 32 hypotheses x 3 roots
constexpr int F_CHUNK = N_MAX / 3;
constexpr int B_SBO = KCHUNKS * 128;         // bytes between 8-template groups of one digit plane
constexpr int POW_STRIDE = 256;    // floats per template row of the incoherent sum (lags on consecutive addresses)
constexpr int M_MAX = 24;          // half frames whose per-template offsets fit the shared-memory table
// Warp layout: warp 0 P builder, warp 1 MMA issuer, then NCG groups of 4 epilogue warps (one per TMEM lane quarter), each
// group owning NC template columns.  Registers per thread follow from the register file (64 K) and the 4-warp allocation
// granularity: 18 warps -> 96, 26 warps -> 72.
__host__ __device__ constexpr int threads(int ncg) { return 64 + 128 * ncg; }
__host__ __device__ constexpr int maxreg(int ncg) { return ncg <= 4 ? 96 : 72; }
constexpr uint32_t TMEM_COLS = 512;
// TMEM map: two "wide" accumulator slots of 192 columns at 0 and 192 (digit planes 0 and 1 side by side, one UTCIMMA with
// N = 2*npad: an instruction costs max(N,128)/2 cycles, so stacking planes is cheaper than issuing them one by one), and
// "narrow" slots of npad columns from 384 (digit plane 2): two when they fit (npad <= 64), else one.
constexpr uint32_t TMEM_WIDE = 192;
constexpr uint32_t TMEM_NARROW0 = 384;
constexpr int RAW_CHUNKS = (2 * NT + KB + 16 + 15 + 15) / 16 + 1;   // 16-byte chunks of raw IQ bytes per tile (+ slack)
// shared-memory map for a chunk padded to npad template columns
constexpr int SMEM_P = 0;                                      // [2 stages][2 variants][P_BYTES]
constexpr int SMEM_B = SMEM_P + 4 * P_BYTES;                   // [3 digits][npad/8][KCHUNKS][8][16] int8
__host__ __device__ constexpr int smem_pow(int npad) { return SMEM_B + 3 * (npad / 8) * B_SBO; }            // [npad][POW_STRIDE] float
__host__ __device__ constexpr int smem_corr(int npad) { return smem_pow(npad) + npad * POW_STRIDE * 4; }    // [2][npad] float
__host__ __device__ constexpr int smem_dsh(int npad) { return smem_corr(npad) + 2 * npad * 4; }             // [M_MAX][npad] int32 byte offsets
__host__ __device__ constexpr int smem_bar(int npad) { return smem_dsh(npad) + M_MAX * npad * 4; }          // 12 mbarriers
__host__ __device__ constexpr int smem_misc(int npad) { return smem_bar(npad) + 16 * 8; }
__host__ __device__ constexpr int smem_raw(int npad) { return smem_misc(npad) + 16 + M_MAX * 4; }
__host__ __device__ constexpr int smem_total(int npad) { return smem_raw(npad) + RAW_CHUNKS * 16 + 16; }
// UTCIMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format S32 (2) bits[4,6);
// a_format / b_format = 1 (signed 8 bit) bits [7,10) / [10,13); K-major A and B; N>>3 bits [17,23); M>>4 bits [24,29)
__host__ __device__ constexpr uint32_t idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
}  // namespace tc

struct TcParams {
  const uint8_t* iq;          // [batch][n_cap][2] raw bytes
  const uint8_t* b_op;        // [3 digits][npad/8][KCHUNKS][8][16] int8 template digits in UMMA core-matrix order
  const int16_t* dsh;         // [n_comb][npad] fold offset of the column's hypothesis minus the chunk minimum
  const int* smin_all;        // [n_comb] min over the chunk's f
  const int* dmax_all;        // [n_comb] max of dsh over the chunk's columns
  const float* corr;          // [2][npad] (C_re row, C_im row)
  float* single_planar;       // [batch][3][n_f][9600]
  uint32_t n_cap, n_f, n_comb, batch;   // n_f = hypotheses in this chunk (<= 32)
  uint32_t f0, n_f_total;     // first hypothesis of the chunk / size of the whole grid
  uint32_t t_tile;            // fold positions per tile
  uint32_t tiles_per_buf;     // ceil(9600 / t_tile)
  float inv_scale;            // 1 / (S * 128)
  long long* prof;            // optional [grid][8] cycle counters (NULL = off)
  uint32_t dbg;               // LCS_TC_DBG builds only: 1 = epilogue releases planes unread, 2 = no fold, 3 = read but no math (timing experiments, wrong results)
};

// ---- small PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), no swizzle
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}
// Variants guarded by an "elected lane" flag so that the issuing warp stays convergent: the compiler then keeps
// descriptors in uniform registers and emits one UTCIMMA per call instead of an elect-and-loop sequence.
__device__ __forceinline__ uint32_t elect_one_flag() {
  uint32_t r;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void umma_i8_g(uint32_t flag, uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(flag)
      : "memory");
}
__device__ __forceinline__ void umma_commit_g(uint32_t flag, uint32_t bar) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar), "r"(flag) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(uint32_t nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// NC = template columns per epilogue warp, NCG = column groups (npad = NC * NCG)
template <int NC, int NCG>
__global__ void __maxnreg__(tc::maxreg(NCG)) xcorr_fold_tc_kernel(const TcParams p) {
  constexpr int NPAD = NCG * NC;
  constexpr int THREADS = tc::threads(NCG), N_EPI_WARPS = 4 * NCG;
  constexpr int B_PLANE = (NPAD / 8) * tc::B_SBO;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3;                   // TMEM lanes 32*quarter .. +31 are accessible to this warp
  const int ewarp = warp - 2;                     // epilogue numbering 0..15
  const int colgrp = ewarp >> 2;                  // which NC of the chunk's template columns
  uint8_t* sP = smem + tc::SMEM_P;
  float* sPow = reinterpret_cast<float*>(smem + tc::smem_pow(NPAD));
  float* sCorr = reinterpret_cast<float*>(smem + tc::smem_corr(NPAD));
  int* sDoff = reinterpret_cast<int*>(smem + tc::smem_dsh(NPAD));
  int* sDmax = reinterpret_cast<int*>(smem + tc::smem_misc(NPAD) + 16);     // [M_MAX] max fold offset per half frame
  const uint32_t bar0 = smem_u32(smem + tc::smem_bar(NPAD));
  // barriers (8 B each): 0,1 p_full[stage]; 2,3 p_empty[stage]; 4,5 wide_full[slot]; 6,7 wide_empty[slot];
  // 8,9 narrow_full[slot]; 10,11 narrow_empty[slot]
  const uint32_t BAR_PFULL = bar0, BAR_PEMPTY = bar0 + 16, BAR_WFULL = bar0 + 32, BAR_WEMPTY = bar0 + 48, BAR_XFULL = bar0 + 64, BAR_XEMPTY = bar0 + 80;
  constexpr int NX = NPAD <= 64 ? 2 : 1;      // narrow slots
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + tc::smem_misc(NPAD));

  // ---- one-time setup ----
  for (int i = tid; i < NPAD * tc::POW_STRIDE; i += THREADS) sPow[i] = 0.f;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.b_op);
    uint4* dst = reinterpret_cast<uint4*>(smem + tc::SMEM_B);
    for (int i = tid; i < 3 * B_PLANE / 16; i += THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < 2 * NPAD; i += THREADS) sCorr[i] = __ldg(p.corr + i);
    for (int i = tid; i < (int)p.n_comb * NPAD; i += THREADS) sDoff[i] = -4 * (int)p.dsh[i];
    for (int i = tid; i < (int)p.n_comb; i += THREADS) sDmax[i] = __ldg(p.dmax_all + i);
  }
  if (tid == 0) {
    mbar_init(BAR_PFULL, 1); mbar_init(BAR_PFULL + 8, 1);
    mbar_init(BAR_PEMPTY, 1); mbar_init(BAR_PEMPTY + 8, 1);            // tcgen05.commit
    for (int i = 0; i < 2; i++) {
      mbar_init(BAR_WFULL + 8 * i, 1); mbar_init(BAR_WEMPTY + 8 * i, N_EPI_WARPS);
      mbar_init(BAR_XFULL + 8 * i, 1); mbar_init(BAR_XEMPTY + 8 * i, N_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(tc::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_async_smem();      // template planes: generic-proxy stores -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t n_items = p.batch * p.tiles_per_buf;
  const uint32_t n_my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t n_tiles = n_my_items * p.n_comb;

  if (warp == 0) {
    // ================= P builder =================
    for (uint32_t tc_i = 0; tc_i < n_tiles; tc_i++) {
      const uint32_t item = blockIdx.x + (tc_i / p.n_comb) * gridDim.x, m = tc_i % p.n_comb;
      const uint32_t b = item / p.tiles_per_buf, i0 = (item % p.tiles_per_buf) * p.t_tile;
      const uint32_t stage = tc_i & 1, use = tc_i >> 1;
      mbar_wait(BAR_PEMPTY + 8 * stage, (use & 1) ^ 1);   // wait until the MMAs that read this stage retired
      if (LCS_TC_DBG && (p.dbg == 4 || p.dbg == 5)) { __syncwarp(); if (lane == 0) mbar_arrive(BAR_PFULL + 8 * stage); continue; }   // timing experiment: no P tiles
      const int64_t z0 = 2 * ((int64_t)i0 + __ldg(p.smin_all + m));          // byte offset of the tile's first lag
      const uint8_t* zb = p.iq + (size_t)b * p.n_cap * 2;
      // Stage the tile's raw bytes (2*NT + KB + alignment slack < 1 KB) with coalesced 128-bit loads, then expand
      // from shared memory: the expansion reads every byte 8 times, global memory only once.
      const int64_t zal = z0 & ~(int64_t)15;
      const int64_t zend = (int64_t)(p.batch - b) * p.n_cap * 2;              // bytes left in the whole allocation
      uint4* raw = reinterpret_cast<uint4*>(smem + tc::smem_raw(NPAD));
      for (int c = lane; c < tc::RAW_CHUNKS; c += 32) {
        const int64_t a = zal + 16 * c;
        raw[c] = (a + 16 <= zend) ? __ldg(reinterpret_cast<const uint4*>(zb + a)) : make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
      }
      __syncwarp();
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw);
      const int zo = (int)(z0 - zal);
      uint4* P1 = reinterpret_cast<uint4*>(sP + (stage * 2 + 0) * tc::P_BYTES);
      uint4* P2 = reinterpret_cast<uint4*>(sP + (stage * 2 + 1) * tc::P_BYTES);
#pragma unroll 2
      for (int row = lane; row < tc::NBLK * 8; row += 32) {     // row = u*8 + r  -> 16 bytes at z0 + 16u + 2r
        const int o = zo + 16 * (row >> 3) + 2 * (row & 7);
        const int ow = o >> 2;
        const bool sh = (o & 3) != 0;
        uint32_t w[5];
#pragma unroll
        for (int i = 0; i < 5; i++) w[i] = rw[ow + i];
        uint32_t x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t v = sh ? __byte_perm(w[i], w[i + 1], 0x5432) : w[i];
          x[i] = v ^ 0x80808080u;                               // (I', Q') = v - 128
          y[i] = __byte_perm(v, 0, 0x2301) ^ 0x7F807F80u;       // (Q', ~I')
        }
        P1[row] = make_uint4(x[0], x[1], x[2], x[3]);
        P2[row] = make_uint4(y[0], y[1], y[2], y[3]);
      }
      __syncwarp();
      fence_async_smem();      // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_PFULL + 8 * stage);
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp walks the pipeline convergently, one elected lane issues ====
    const uint32_t sP_addr = smem_u32(sP), sB_addr = smem_u32(smem + tc::SMEM_B);
    const uint32_t flag = elect_one_flag();
    constexpr uint32_t IDESC_W = tc::idesc(2 * NPAD), IDESC_X = tc::idesc(NPAD);
    // descriptors advance by adding to the 14-bit (address >> 4) field: +16 per 256-byte K step
    const uint64_t bw_desc = make_desc(sB_addr, 128, tc::B_SBO);                  // digit planes 0|1: 2*NPAD template rows
    const uint64_t bx_desc = make_desc(sB_addr + 2 * B_PLANE, 128, tc::B_SBO);    // digit plane 2
    uint32_t st = 0;   // running sub-tile counter: every sub-tile uses wide slot v for part v (re/im) once
    long long t_pwait = 0, t_ewait = 0, t_xwait = 0, t_start = TC_CLK();
    for (uint32_t tc_i = 0; tc_i < n_tiles; tc_i++) {
      const uint32_t stage = tc_i & 1, use = tc_i >> 1;
      long long c0 = TC_CLK();
      mbar_wait(BAR_PFULL + 8 * stage, use & 1);
      t_pwait += TC_CLK() - c0;
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < tc::NSUB; q++, st++) {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          const uint64_t a_desc = make_desc(sP_addr + (stage * 2 + v) * tc::P_BYTES + q * (tc::NSUBL / 8) * 128, 128, 128);
          // wide job: slot v, used once per sub-tile
          c0 = TC_CLK();
          mbar_wait(BAR_WEMPTY + 8 * v, (st & 1) ^ 1);
          t_ewait += TC_CLK() - c0;
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < tc::KSTEPS; s++)
            umma_i8_g(flag, tmem_base + v * tc::TMEM_WIDE, a_desc + (uint64_t)(s * 16), bw_desc + (uint64_t)(s * 16), IDESC_W, s > 0);
          umma_commit_g(flag, BAR_WFULL + 8 * v);
          // narrow job: NX == 2: slot v once per sub-tile; NX == 1: slot 0 twice per sub-tile
          const uint32_t xs = NX == 2 ? v : 0;
          const uint32_t xpar = NX == 2 ? (st & 1) : (uint32_t)v;
          c0 = TC_CLK();
          mbar_wait(BAR_XEMPTY + 8 * xs, xpar ^ 1);
          t_xwait += TC_CLK() - c0;
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < tc::KSTEPS; s++)
            umma_i8_g(flag, tmem_base + tc::TMEM_NARROW0 + xs * NPAD, a_desc + (uint64_t)(s * 16), bx_desc + (uint64_t)(s * 16), IDESC_X, s > 0);
          umma_commit_g(flag, BAR_XFULL + 8 * xs);
        }
      }
      umma_commit_g(flag, BAR_PEMPTY + 8 * stage);                  // P stage free again
    }
    if (LCS_TC_PROFILE && p.prof && lane == 0) {
      p.prof[blockIdx.x * 8 + 0] = TC_CLK() - t_start;
      p.prof[blockIdx.x * 8 + 1] = t_pwait;
      p.prof[blockIdx.x * 8 + 2] = t_ewait;
      p.prof[blockIdx.x * 8 + 3] = t_xwait;
    }
  } else {
    // ================= epilogue: TMEM -> |xc|^2 -> fold =================
    const int L = quarter * 32 + lane;              // lag row of the sub-tile
    const int col0 = colgrp * NC;                   // first template column of this warp
    const uint32_t n_templ = 3 * p.n_f;
    const float inv2s = p.inv_scale * p.inv_scale;  // inv_scale is a power of two: scaling commutes with the roundings
    char* myPowB = reinterpret_cast<char*>(sPow + col0 * tc::POW_STRIDE + L);
    const float* cre = sCorr + col0;
    const float* cim = sCorr + NPAD + col0;
    const uint32_t dbg = LCS_TC_DBG ? p.dbg : 0;
    long long t_fwait = 0, t_ld = 0, e_start = TC_CLK();
    // Wide slot v (re/im) holds digit planes 0 and 1 of the current sub-tile, the narrow slot digit plane 2.  This warp's
    // 32 lags x NC templates go to registers, then the slot is released.  Parities: the wide slots and (NX == 2) the narrow
    // slots are used once per sub-tile, the single narrow slot (NX == 1) twice.
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + col0;
    auto ld_cols = [&](uint32_t src, int (&dst)[NC]) {
      if (NC >= 16) tmem_ld16(src, *reinterpret_cast<int(*)[16]>(&dst[0]));
      if (NC == 8 || NC == 24) tmem_ld8(src + (NC - 8), *reinterpret_cast<int(*)[8]>(&dst[NC - 8]));
    };
    long long t_w4[4] = {0, 0, 0, 0};
    auto drain_wide = [&](int (&d0)[NC], int (&d1)[NC], const int v, const uint32_t st_par) {
      long long c0 = TC_CLK();
      mbar_wait(BAR_WFULL + 8 * v, st_par);
      long long c1 = TC_CLK();
      t_fwait += c1 - c0;
      t_w4[2 * v] += c1 - c0;
      tc_fence_after();
      if (dbg != 1 && dbg != 5) {
        ld_cols(lane_base + v * tc::TMEM_WIDE, d0);
        ld_cols(lane_base + v * tc::TMEM_WIDE + NPAD, d1);
        tmem_ld_wait();
      }
      t_ld += TC_CLK() - c1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_WEMPTY + 8 * v);
    };
    auto drain_narrow = [&](int (&d2)[NC], const int v, const uint32_t st_par) {
      const uint32_t xs = NX == 2 ? v : 0;
      long long c0 = TC_CLK();
      mbar_wait(BAR_XFULL + 8 * xs, NX == 2 ? st_par : (uint32_t)v);
      long long c1 = TC_CLK();
      t_fwait += c1 - c0;
      t_w4[2 * v + 1] += c1 - c0;
      tc_fence_after();
      if (dbg != 1 && dbg != 5) {
        ld_cols(lane_base + tc::TMEM_NARROW0 + xs * NPAD, d2);
        tmem_ld_wait();
      }
      t_ld += TC_CLK() - c1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_XEMPTY + 8 * xs);
    };
    uint32_t st_par = 0;        // parity of the running sub-tile counter
    for (uint32_t it = 0; it < n_my_items; it++) {
      const uint32_t item = blockIdx.x + it * gridDim.x;
      const uint32_t b = item / p.tiles_per_buf, i0 = (item % p.tiles_per_buf) * p.t_tile;
      for (uint32_t m = 0; m < p.n_comb; m++) {
        const int dmax = sDmax[m];
        const int* doff = sDoff + m * NPAD + col0;        // -4 * (fold offset of the column - chunk minimum), bytes
#pragma unroll
        for (int q = 0; q < tc::NSUB; q++, st_par ^= 1) {
          // Recombine the three digit planes: t = a0*256 + a1 (int32, exact), value = float(t)*256 + float(a2).
          int t[NC], a[NC];
          float rr[NC];
          if (dbg == 1 || dbg == 3 || dbg == 5) {
            for (int v = 0; v < 2; v++) { drain_wide(t, a, v, st_par); drain_narrow(a, v, st_par); }
            if (dbg == 3 && t[0] == 0x7fffffff) sPow[0] = 1.f;
            continue;
          }
          // ---- real part ----
          drain_wide(t, a, 0, st_par);
#pragma unroll
          for (int c = 0; c < NC; c++) t[c] = t[c] * 256 + a[c];
          drain_narrow(a, 0, st_par);
#pragma unroll
          for (int c = 0; c < NC; c += 4) {
            const float4 k = *reinterpret_cast<const float4*>(cre + c);
            const float kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float x = __fadd_rn(__fmaf_rn((float)t[c + e], 256.f, (float)a[c + e]), kk[e]);
              rr[c + e] = __fmul_rn(x, x);
            }
          }
          // ---- imaginary part ----
          drain_wide(t, a, 1, st_par);
#pragma unroll
          for (int c = 0; c < NC; c++) t[c] = t[c] * 256 + a[c];
          drain_narrow(a, 1, st_par);
          // |xc|^2 = re^2 + im^2 (searcher.cpp:300), in the integer scale of the templates; the power-of-two scale
          // factor is applied when the tile is written out.  rr[] becomes the tile's contribution to the fold.
#pragma unroll
          for (int c = 0; c < NC; c += 4) {
            const float4 k = *reinterpret_cast<const float4*>(cim + c);
            const float kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float x = __fadd_rn(__fmaf_rn((float)t[c + e], 256.f, (float)a[c + e]), kk[e]);
              rr[c + e] = __fmaf_rn(x, x, rr[c + e]);
            }
          }
          // fold: all loads of the read-modify-write first (t[] is dead, its registers hold the addresses), then add + store
          if (dbg == 2) {
            if (rr[0] + rr[NC - 1] == 1.2345f) sPow[1] = 1.f;
            continue;
          }
#pragma unroll
          for (int c = 0; c < NC; c += 4) {
            const int4 d4 = *reinterpret_cast<const int4*>(doff + c);
            t[c] = d4.x; t[c + 1] = d4.y; t[c + 2] = d4.z; t[c + 3] = d4.w;
          }
          // this lane's lag inside the tile (before the column's fold offset) is q*128 + L
          const bool inside = (q * tc::NSUBL + quarter * 32 - dmax >= 0) && (q * tc::NSUBL + quarter * 32 + 32 <= (int)p.t_tile);   // warp-uniform
          float cur[NC];
          if (inside) {
#pragma unroll
            for (int c = 0; c < NC; c++) cur[c] = *reinterpret_cast<const float*>(myPowB + t[c] + (c * tc::POW_STRIDE + q * tc::NSUBL) * 4);
#pragma unroll
            for (int c = 0; c < NC; c++) *reinterpret_cast<float*>(myPowB + t[c] + (c * tc::POW_STRIDE + q * tc::NSUBL) * 4) = __fadd_rn(cur[c], rr[c]);
          } else {
#pragma unroll
            for (int c = 0; c < NC; c++)
              if ((unsigned)(q * tc::NSUBL * 4 + L * 4 + t[c]) < p.t_tile * 4) {
                float* dst = reinterpret_cast<float*>(myPowB + t[c] + (c * tc::POW_STRIDE + q * tc::NSUBL) * 4);
                *dst = __fadd_rn(*dst, rr[c]);
              }
          }
        }
      }
      // ---- item done: write xc_incoherent_single rows (coalesced), reset the accumulators ----
      epi_bar(32 * N_EPI_WARPS);
      const float ncf = (float)p.n_comb;
      for (uint32_t row = ewarp; row < n_templ; row += N_EPI_WARPS) {
        const uint32_t rf = row / 3, rt = row % 3;
        float* dst = p.single_planar + (((size_t)b * 3 + rt) * p.n_f_total + p.f0 + rf) * LCS_N_FOLD + i0;
        float* src = sPow + row * tc::POW_STRIDE;
        for (uint32_t i = lane; i < p.t_tile; i += 32) {
          if (i0 + i < LCS_N_FOLD) dst[i] = __fdiv_rn(__fmul_rn(src[i], inv2s), ncf);   // searcher.cpp:304
          src[i] = 0.f;
        }
      }
      epi_bar(32 * N_EPI_WARPS);
    }
    if (LCS_TC_PROFILE && p.prof && lane == 0 && ewarp == 0) {
      p.prof[blockIdx.x * 8 + 4] = TC_CLK() - e_start;
      p.prof[blockIdx.x * 8 + 5] = t_fwait;
      p.prof[blockIdx.x * 8 + 6] = t_ld;
      for (int k = 0; k < 4; k++) p.prof[148 * 8 + blockIdx.x * 4 + k] = t_w4[k];
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tc::TMEM_COLS));
  }
}

// =============================================================================================
// Host side
// =============================================================================================
static uint32_t tc_npad(uint32_t n_f_chunk) { return (3 * n_f_chunk + 31) / 32 * 32; }

lcs_status tc_plan_setup(lcs_xcorr_plan* p) {
  p->tc_ready = false;
  const XcorrGeom& g = p->geom;
  // Hypotheses are processed in chunks of <= 32 (3*32 = 96 template columns of the UMMA N dimension).
  const uint32_t n_chunks = (g.n_f + tc::F_CHUNK - 1) / tc::F_CHUNK;
  const uint32_t chunk = (g.n_f + n_chunks - 1) / n_chunks;
  if (n_chunks > 8 || g.n_comb_xc > (uint32_t)tc::M_MAX) return LCS_OK;
  // fold-offset spread inside a chunk decides how many fold positions a 256-lag tile yields
  std::vector<int> smin((size_t)n_chunks * g.n_comb_xc), dmax((size_t)n_chunks * g.n_comb_xc);
  int spread = 0;
  for (uint32_t c = 0; c < n_chunks; c++)
    for (uint32_t m = 0; m < g.n_comb_xc; m++) {
      int lo = INT32_MAX, hi = INT32_MIN;
      for (uint32_t f = c * chunk; f < std::min(g.n_f, (c + 1) * chunk); f++) {
        lo = std::min(lo, p->h_soff[(size_t)m * g.n_f + f]);
        hi = std::max(hi, p->h_soff[(size_t)m * g.n_f + f]);
      }
      smin[(size_t)c * g.n_comb_xc + m] = lo;
      dmax[(size_t)c * g.n_comb_xc + m] = hi - lo;
      spread = std::max(spread, hi - lo);
    }
  if (spread > tc::NT - 64) return LCS_OK;         // grid too sparse for this tiling: the FP32 kernel handles it
  const int t_tile = tc::NT - spread;

  // scale: power of two with |W*S| <= 127*65536 + 127*256 + 127
  double maxabs = 0;
  for (const cd& w : p->h_w) maxabs = std::max(maxabs, std::max(std::fabs(w.real()), std::fabs(w.imag())));
  const double limit = 127.0 * 65536 + 127 * 256 + 127;
  int e = (int)std::floor(std::log2(limit / maxabs));
  while (std::ldexp(maxabs, e) > limit) e--;
  const double S = std::ldexp(1.0, e);

  // per chunk: digit planes in core-matrix order, corrections, fold-offset table (all sized for N_MAX columns)
  constexpr size_t B_CHUNK_BYTES = (size_t)3 * (tc::N_MAX / 8) * tc::B_SBO;
  std::vector<uint8_t> b_op(n_chunks * B_CHUNK_BYTES, 0);
  std::vector<float> corr((size_t)n_chunks * 2 * tc::N_MAX, 0.f);
  std::vector<int16_t> dsh((size_t)n_chunks * tc::M_MAX * tc::N_MAX, 0);
  for (uint32_t f = 0; f < g.n_f; f++) {
    const uint32_t c = f / chunk, fl = f - c * chunk;
    const uint32_t npad = tc_npad(std::min(chunk, g.n_f - c * chunk));
    const size_t plane = (size_t)(npad / 8) * tc::B_SBO;
    uint8_t* bc = b_op.data() + c * B_CHUNK_BYTES;
    auto put = [&](int col, int k, long long wint) {
      // balanced base-256 digits: wint = 65536 d0 + 256 d1 + d2, d1,d2 in [-128,127]
      long long d2 = ((wint % 256) + 256) % 256; if (d2 > 127) d2 -= 256;
      long long r1 = (wint - d2) / 256;
      long long d1 = ((r1 % 256) + 256) % 256; if (d1 > 127) d1 -= 256;
      long long d0 = (r1 - d1) / 256;
      const long long dig[3] = {d0, d1, d2};
      // core matrix (column group col/8, K chunk k/16): 8 rows of 16 bytes
      const size_t off = (size_t)(col / 8) * tc::B_SBO + (size_t)(k / 16) * 128 + (size_t)(col % 8) * 16 + (k % 16);
      for (int j = 0; j < 3; j++) bc[j * plane + off] = (uint8_t)(int8_t)dig[j];
    };
    for (int t = 0; t < 3; t++) {
      const int col = (int)fl * 3 + t;
      long long sum_all = 0, sum_even = 0;
      for (int tap = 0; tap < 137; tap++) {
        const cd w = p->h_w[((size_t)f * 3 + t) * 137 + tap];
        const long long wr = std::llrint(w.real() * S), wi = std::llrint(w.imag() * S);
        put(col, 2 * tap, wr);        // multiplies the I byte
        put(col, 2 * tap + 1, -wi);   // multiplies the Q byte (re) / ~I byte (im)
        sum_all += wr - wi;
        sum_even += wr;
      }
      corr[(size_t)c * 2 * tc::N_MAX + col] = (float)sum_all;           // x = x'+1 :  + sum_j a[j]
      corr[(size_t)c * 2 * tc::N_MAX + npad + col] = (float)sum_even;   // (Q', ~I') stream:  + sum_m a[2m]
      for (uint32_t m = 0; m < g.n_comb_xc; m++)
        dsh[(size_t)c * tc::M_MAX * tc::N_MAX + (size_t)m * npad + col] = (int16_t)(p->h_soff[(size_t)m * g.n_f + f] - smin[(size_t)c * g.n_comb_xc + m]);
    }
  }
  lcs_ctx* ctx = p->ctx;
  std::vector<int> meta(smin);
  meta.insert(meta.end(), dmax.begin(), dmax.end());
  LCS_CUDA(ctx, p->d_tc_a.alloc(b_op.size()));
  LCS_CUDA(ctx, p->d_tc_meta.alloc(meta.size()));
  LCS_CUDA(ctx, p->d_tc_scale.alloc(corr.size()));
  LCS_CUDA(ctx, p->d_tc_dsh.alloc(dsh.size()));
  LCS_CUDA(ctx, cudaMemcpy(p->d_tc_a.p, b_op.data(), b_op.size(), cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, cudaMemcpy(p->d_tc_meta.p, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, cudaMemcpy(p->d_tc_scale.p, corr.data(), corr.size() * 4, cudaMemcpyHostToDevice));
  LCS_CUDA(ctx, cudaMemcpy(p->d_tc_dsh.p, dsh.data(), dsh.size() * 2, cudaMemcpyHostToDevice));
  p->tc_params[0] = t_tile;
  p->tc_params[1] = (LCS_N_FOLD + t_tile - 1) / t_tile;
  float inv = (float)(1.0 / (S * 128.0));
  std::memcpy(&p->tc_params[2], &inv, 4);
  p->tc_params[3] = (int)n_chunks;
  p->tc_params[4] = (int)chunk;
  LCS_CUDA(ctx, cudaFuncSetAttribute(xcorr_fold_tc_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::smem_total(32)));
  LCS_CUDA(ctx, cudaFuncSetAttribute(xcorr_fold_tc_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::smem_total(64)));
  LCS_CUDA(ctx, cudaFuncSetAttribute(xcorr_fold_tc_kernel<24, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::smem_total(96)));
  LCS_CUDA(ctx, cudaFuncSetAttribute(xcorr_fold_tc_kernel<16, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::smem_total(96)));
  p->tc_ready = true;
  return LCS_OK;
}

// LCS_TC_PROF=1 : per-CTA cycle counters of the pipeline stages are collected and printed at exit (debug aid).
static long long* g_prof = nullptr;
static long long* tc_prof_buffer() {
  static int on = -1;
  if (on < 0) on = std::getenv("LCS_TC_PROF") ? 1 : 0;
  if (!on) return nullptr;
  if (!g_prof) { cudaMalloc((void**)&g_prof, 148 * 12 * 8); cudaMemset(g_prof, 0, 148 * 12 * 8); }
  return g_prof;
}
void tc_prof_dump() {
  if (!g_prof) return;
  std::vector<long long> h(148 * 12);
  cudaDeviceSynchronize();
  cudaMemcpy(h.data(), g_prof, h.size() * 8, cudaMemcpyDeviceToHost);
  for (int b : {0, 1, 73, 147})
    std::printf("[tc prof] cta %3d: mma total %lld  wait_P %lld  wait_wide_empty %lld  wait_narrow_empty %lld | epi(warp 2) total %lld  wait_full %lld (W.re %lld X.re %lld W.im %lld X.im %lld)  tmem_ld %lld\n", b, h[b * 8], h[b * 8 + 1],
                h[b * 8 + 2], h[b * 8 + 3], h[b * 8 + 4], h[b * 8 + 5], h[148 * 8 + b * 4], h[148 * 8 + b * 4 + 1], h[148 * 8 + b * 4 + 2], h[148 * 8 + b * 4 + 3], h[b * 8 + 6]);
}

int launch_xcorr_fold_tc(lcs_xcorr_plan* p, const void* d_iq_cu8, uint32_t batch, float* d_single_planar, cudaStream_t st) {
  const uint32_t n_chunks = (uint32_t)p->tc_params[3], chunk = (uint32_t)p->tc_params[4];
  constexpr size_t B_CHUNK_BYTES = (size_t)3 * (tc::N_MAX / 8) * tc::B_SBO;
  for (uint32_t c = 0; c < n_chunks; c++) {
    TcParams q;
    q.iq = reinterpret_cast<const uint8_t*>(d_iq_cu8);
    q.b_op = p->d_tc_a.p + c * B_CHUNK_BYTES;
    q.dsh = p->d_tc_dsh.p + (size_t)c * tc::M_MAX * tc::N_MAX;
    q.smin_all = p->d_tc_meta.p + (size_t)c * p->geom.n_comb_xc;
    q.dmax_all = p->d_tc_meta.p + (size_t)(n_chunks + c) * p->geom.n_comb_xc;
    q.corr = p->d_tc_scale.p + (size_t)c * 2 * tc::N_MAX;
    q.single_planar = d_single_planar;
    q.n_cap = p->geom.n_cap;
    q.f0 = c * chunk;
    q.n_f = std::min(chunk, p->geom.n_f - q.f0);
    q.n_f_total = p->geom.n_f;
    q.n_comb = p->geom.n_comb_xc;
    q.batch = batch;
    q.t_tile = (uint32_t)p->tc_params[0];
    q.tiles_per_buf = (uint32_t)p->tc_params[1];
    std::memcpy(&q.inv_scale, &p->tc_params[2], 4);
    q.prof = tc_prof_buffer();
    q.dbg = std::getenv("LCS_TC_DBG") ? (uint32_t)std::atoi(std::getenv("LCS_TC_DBG")) : 0;
    const uint32_t n_items = batch * q.tiles_per_buf;
    const uint32_t grid = std::min<uint32_t>((uint32_t)p->ctx->n_sm, n_items);
    const uint32_t npad = tc_npad(q.n_f);
    // 96 columns: 6 groups of 16 (26 warps, 72 registers) measured 3.5 % faster than 4 groups of 24 (18 warps, 96 registers);
    // LCS_TC_LAYOUT=4 selects the latter for comparison
    static const int layout6 = std::getenv("LCS_TC_LAYOUT") ? std::atoi(std::getenv("LCS_TC_LAYOUT")) : 6;
    if (npad == 32) xcorr_fold_tc_kernel<8, 4><<<grid, tc::threads(4), tc::smem_total(32), st>>>(q);
    else if (npad == 64) xcorr_fold_tc_kernel<16, 4><<<grid, tc::threads(4), tc::smem_total(64), st>>>(q);
    else if (layout6 == 6) xcorr_fold_tc_kernel<16, 6><<<grid, tc::threads(6), tc::smem_total(96), st>>>(q);
    else xcorr_fold_tc_kernel<24, 4><<<grid, tc::threads(4), tc::smem_total(96), st>>>(q);
  }
  return (int)n_chunks;
}

}  // namespace lcs
