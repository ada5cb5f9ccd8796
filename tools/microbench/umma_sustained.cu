// Microbenchmark: what the int8 tensor pipe of a B200 sustains with ALL SMs busy, in the shapes the correlator issues
// (tcgen05.mma kind::i8, M=128, N=144, K=32 B, operands from shared memory with random data, accumulators rotating over
// three TMEM slots, one commit per 9 instructions).  Reports int8 POP/s for a short burst and for a seconds-long run
// (power-capped steady state) together with the SM clock derived from clock64() against wall time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_sustained umma_sustained.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int N = 144, KSTEPS = 9, NSLOT = 3;

__global__ void sustained(int jobs, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[NSLOT];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  uint32_t x = 1234567u + blockIdx.x * 7919u + tid;
  for (int i = tid; i < 100 * 1024; i += blockDim.x) { x = x * 1664525u + 1013904223u; smem[i] = (uint8_t)(x >> 24); }
  if (tid == 0)
    for (int i = 0; i < NSLOT; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16 * 1024);
    const long long t0 = clock64();
    for (int j = 0; j < jobs; j++) {
      const int slot = j % NSLOT;
      if (j >= NSLOT) mbar_wait(smem_u32(&bars[slot]), ((j / NSLOT) - 1) & 1);     // the previous job in this slot has retired
      const uint64_t da = make_desc(sa + (j & 3) * 2048, 128, 128);                 // Hankel-style A (LBO = SBO = 128 B)
      const uint64_t db = make_desc(sb + (j & 1) * (N / 8) * 2304, 128, 2304);      // K-major template planes
      for (int s = 0; s < KSTEPS; s++) mma_i8(tb + slot * N, da + s * 16, db + s * 16, idesc, s > 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[slot])) : "memory");
    }
    for (int j = jobs - NSLOT; j < jobs; j++)
      if (j >= 0) mbar_wait(smem_u32(&bars[j % NSLOT]), (j / NSLOT) & 1);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, n_sm * 8);
  cudaFuncSetAttribute(sustained, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double ops_per_job = 2.0 * 128 * N * 32 * KSTEPS;
  sustained<<<n_sm, 128, 100 * 1024>>>(2000, d);      // warm-up
  cudaDeviceSynchronize();
  for (int jobs : {20000, 200000, 2000000, 6000000}) {
    cudaEventRecord(e0);
    sustained<<<n_sm, 128, 100 * 1024>>>(jobs, d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c[256] = {0};
    cudaMemcpy(c, d, n_sm * 8, cudaMemcpyDeviceToHost);
    long long cmax = 0;
    for (int i = 0; i < n_sm; i++) cmax = c[i] > cmax ? c[i] : cmax;
    printf("%d SMs x %8d jobs of 9 x (M=128,N=%d,K=32B) int8: %9.2f ms  %.3f int8 POP/s  %.1f cycles/MMA  SM clock %.0f MHz  (%s)\n", n_sm, jobs, N,
           ms, ops_per_job * jobs * n_sm / (ms * 1e-3) / 1e15, (double)cmax / jobs / KSTEPS, cmax / (ms * 1e-3) / 1e6, cudaGetErrorString(err));
  }
  return 0;
}
