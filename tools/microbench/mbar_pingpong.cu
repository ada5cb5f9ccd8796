// Microbenchmark: latency of an mbarrier hand-off between two warps of a CTA (cycles per round trip), with the
// waiting side using (0) mbarrier.try_wait (hardware-suspended wait) or (1) a mbarrier.test_wait spin loop, and
// the latency from the last tcgen05.mma of a chain to the waiter seeing its tcgen05.commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_pingpong mbar_pingpong.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int SPIN>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  if (SPIN) {
    uint32_t done = 0;
    while (!done) asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } else {
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(bar), "r"(parity) : "memory");
  }
}
template <int SPIN>
__global__ void pingpong(int iters, int nwait, long long* out) {
  __shared__ uint64_t bars[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[1])), "r"(nwait));
  }
  __syncthreads();
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  long long t0 = clock64();
  if (warp == 0) {           // "producer": signals b0, waits for all nwait consumers on b1
    for (int i = 0; i < iters; i++) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");
      wait<SPIN>(b1, i & 1);
    }
    if (lane == 0) out[0] = clock64() - t0;
  } else if (warp <= nwait) {  // consumers: wait for b0, then arrive on b1
    for (int i = 0; i < iters; i++) {
      wait<SPIN>(b0, i & 1);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
    }
  }
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int spin = 0; spin < 2; spin++)
    for (int nwait : {1, 4, 16}) {
      const int iters = 20000;
      if (spin) pingpong<1><<<1, 32 * 17>>>(iters, nwait, d); else pingpong<0><<<1, 32 * 17>>>(iters, nwait, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("%-10s consumers=%2d : %7.1f cycles per round trip (%s)\n", spin ? "test_wait" : "try_wait", nwait, (double)c / iters, cudaGetErrorString(e));
    }
  return 0;
}
