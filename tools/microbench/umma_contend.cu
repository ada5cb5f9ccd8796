// Microbenchmark: tcgen05.mma (.ss and .ts, kind::i8) rate while 16 other warps hammer shared memory with LDS/FADD/STS
// (does operand fetch from shared memory compete with an epilogue's read-modify-write traffic?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_contend umma_contend.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
template <int KIND, int ATMEM>
__device__ __forceinline__ void mma(uint32_t d, uint32_t at, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0) {
    if (ATMEM) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(at), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  } else {
    if (ATMEM) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(at), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  }
}

template <int ATMEM>
__global__ void bench(int N, int iters, int delay, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 160 * 1024; i += blockDim.x) smem[i] = (uint8_t)(i * 7);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); stop = 0; }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  if (warp == 0) {
    if (lane == 0) {
      uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
      long long t0 = clock64();
      for (int i = 0; i < iters; i++) {
        const uint64_t da = make_desc(sa + (i & 7) * 256, 128, 128), db = make_desc(sb + (i & 7) * 256, 128, 2304);
        mma<0, ATMEM>(tb + 256, tb + (i & 7) * 8, da, db, idesc, i > 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
      long long t1 = clock64();
      out[0] = t1 - t0;
      stop = 1;
    }
  } else if (delay >= 0 && (warp & 3) != 0) {   // keep the issuing warp's scheduler free: the arbiter favours high warp ids
    // background: 24 x (LDS, FADD, STS), lane-consecutive addresses (1 wavefront each), then `delay` dependent FMAs
    float* base = reinterpret_cast<float*>(smem + 64 * 1024) + warp * 768 + lane;
    float acc = lane;
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int c = 0; c < 24; c++) base[c * 32] = base[c * 32] + acc;
      for (int d = 0; d < delay; d++) acc = fmaf(acc, 1.0001f, 0.5f);
      n++;
    }
    if (lane == 0) out[warp] = n;
    if (acc == 123.456f) out[40] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
template <int ATMEM>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(bench<ATMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int N : {96, 128})
    for (int delay : {-1, 0, 64, 256, 1024}) {
      const int iters = 20000;
      cudaMemset(d, 0, 64 * 8);
      bench<ATMEM><<<1, 17 * 32, 160 * 1024>>>(N, iters, delay, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[64];
      cudaMemcpy(c, d, 64 * 8, cudaMemcpyDeviceToHost);
      long long bg = 0;
      for (int w = 1; w <= 16; w++) bg += c[w];   // 12 background warps (w % 4 != 0)
      printf("%-10s N=%3d bg delay %5d : %6.1f cycles/MMA ; background %.3f LDS+STS wavefronts/cycle (%s)\n", name, N, delay, (double)c[0] / iters,
             (double)bg * 48 / (double)c[0], cudaGetErrorString(e));
    }
  cudaFree(d);
}
int main() {
  run<0>("i8 .ss");
  run<1>("i8 .ts");
  return 0;
}
