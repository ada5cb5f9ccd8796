// Microbenchmark: cycles per tcgen05.mma instruction vs N, kind (i8 / f16-bf16) and A source (smem / tmem).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run on a B200.
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
template <int KIND, int ATMEM>
__global__ void bench(int N, int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * 1024; i += blockDim.x) smem[i] = (uint8_t)(i * 7);
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    // i8: c S32(2), a,b signed(1); f16: c F32(1), a,b BF16(1)
    uint32_t idesc = KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : ((1u << 4) | (1u << 7) | (1u << 10));
    idesc |= ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
      const uint64_t da = make_desc(sa + (i & 7) * 256, 128, 2304), db = make_desc(sb + (i & 7) * 256, 128, 128);
      mma<KIND, ATMEM>(tb + 256, tb + (i & 7) * 8, da, db, idesc, i > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
template <int KIND, int ATMEM>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench<KIND, ATMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {16, 32, 64, 96, 128, 192, 256}) {
    if (N > 256) continue;
    const int iters = 2000;
    bench<KIND, ATMEM><<<1, 128, 64 * 1024>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%-22s M=128 N=%3d K=32B : %7.1f cycles/MMA  (%s)\n", name, N, (double)c / iters, cudaGetErrorString(e));
  }
  cudaFree(d);
}

// Alternating shapes as in the correlator: 9 x (N=192) then 9 x (N=96), A = overlapping "Hankel" descriptor (LBO=SBO=128),
// B = K-major core-matrix planes (LBO 128, SBO 2304); separate accumulators.
__global__ void bench_mix(int iters, int mode, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 120 * 1024; i += blockDim.x) smem[i] = (uint8_t)(i * 7);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3)) : "memory");   // phase 0 complete: waits on parity 0 pass
  }
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
    const uint32_t base = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t idw = base | ((192u >> 3) << 17), idx = base | ((96u >> 3) << 17);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
      const uint64_t da = make_desc(sa + (i & 1) * 2048, 128, 128);
      const uint64_t dbw = make_desc(sb, 128, 2304), dbx = make_desc(sb + 24 * 2304, 128, 2304);
      if (mode != 2) {
        for (int s = 0; s < 9; s++) mma<0, 0>(tb + (i & 1) * 192, 0, da + s * 16, dbw + s * 16, idw, s > 0);
        if (mode >= 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      if (mode != 1) {
        for (int s = 0; s < 9; s++) mma<0, 0>(tb + 384, 0, da + s * 16, dbx + s * 16, idx, s > 0);
        if (mode >= 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      if (mode == 4) {   // as the kernel does: wait for the group before last to complete before issuing more
        asm volatile("{.reg .pred p; W2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D2; bra W2; D2: }" ::"r"(smem_u32(&bar3)), "r"(0) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
    out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
void run_mix() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const char* names[5] = {"9x N=192 + 9x N=96 alternating", "9x N=192 only", "9x N=96 only", "alternating + commit per group", "alternating + commit + mbarrier poll"};
  for (int mode = 0; mode < 5; mode++) {
    const int iters = 500;
    bench_mix<<<1, 128, 120 * 1024>>>(iters, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("%-34s : %8.1f cycles per group (%s)\n", names[mode], (double)c / iters, cudaGetErrorString(e));
  }
  cudaFree(d);
}

int main() {
  run_mix();
  run<0, 0>("i8  A=smem");
  run<0, 1>("i8  A=tmem");
  run<1, 0>("bf16 A=smem");
  run<1, 1>("bf16 A=tmem");
  return 0;
}
