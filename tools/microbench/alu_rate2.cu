// Microbenchmark: issue cost (cycles per warp-instruction per SMSP) of the epilogue's instruction mix, alone and paired,
// 4 warps per SMSP, 8 independent chains per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate2 alu_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
template <int OP>
__global__ void k(int* out, int iters, long long* cyc) {
  int a[8];
  float f[8];
  unsigned long long g[8];
  __shared__ float sm[512 * 9];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; f[i] = a[i]; g[i] = 0x3f8000003f800000ull + i; sm[threadIdx.x + 512 * i] = i; }
  __syncthreads();
  float* sp = sm + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#define I2F(i) asm volatile("{.reg .f32 t; cvt.rn.f32.s32 t, %0; mov.b32 %0, t;}" : "+r"(a[i]));
#define IMAD(i) asm volatile("mad.lo.s32 %0, %0, 256, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
#define IADD(i) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
#define FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
#define FMUL(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
#define FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
#define FFMA2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(g[i]) : "l"(g[(i + 1) & 7]), "l"(g[(i + 2) & 7]));
#define FADD2(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(g[i]) : "l"(g[(i + 1) & 7]));
#define FMUL2(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(g[i]) : "l"(g[(i + 1) & 7]));
#define LDS(i) asm volatile("{.reg .f32 t; ld.shared.f32 t, [%1]; add.rn.f32 %0, %0, t;}" : "+f"(f[i]) : "r"((unsigned)__cvta_generic_to_shared(sp + 512 * i)));
#define STS(i) asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(sp + 512 * i)), "f"(f[i]));
#define I2F_FFMA(i) I2F(i) FFMA(i)
#define I2F_IMAD(i) I2F(i) IMAD(i)
#define I2F_FFMA2(i) I2F(i) FFMA2(i)
#define IMAD_FADD(i) IMAD(i) FADD(i)
#define EPI(i) IMAD(i) I2F(i) FFMA(i) FADD(i) FMUL(i)
#define FOLD(i) LDS(i) STS(i)
    if (OP == 0) { REP8(I2F) }
    if (OP == 1) { REP8(IMAD) }
    if (OP == 2) { REP8(IADD) }
    if (OP == 3) { REP8(FADD) }
    if (OP == 4) { REP8(FMUL) }
    if (OP == 5) { REP8(FFMA) }
    if (OP == 6) { REP8(FFMA2) }
    if (OP == 7) { REP8(FADD2) }
    if (OP == 8) { REP8(FMUL2) }
    if (OP == 9) { REP8(LDS) }
    if (OP == 10) { REP8(STS) }
    if (OP == 11) { REP8(I2F_FFMA) }
    if (OP == 12) { REP8(I2F_IMAD) }
    if (OP == 13) { REP8(I2F_FFMA2) }
    if (OP == 14) { REP8(IMAD_FADD) }
    if (OP == 15) { REP8(FOLD) }
    if (OP == 16) { REP8(EPI) }
  }
  long long t1 = clock64();
  int s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + (int)f[i] + (int)g[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int per) {
  int* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  k<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-14s %6.3f cycles per warp-instruction per SMSP (%s)\n", name, (double)c / (4.0 * iters * 8 * per), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("I2F", 1); run<1>("IMAD", 1); run<2>("IADD", 1); run<3>("FADD", 1); run<4>("FMUL", 1); run<5>("FFMA", 1);
  run<6>("FFMA2", 1); run<7>("FADD2", 1); run<8>("FMUL2", 1); run<9>("LDS+FADD", 2); run<10>("STS", 1);
  run<11>("I2F+FFMA", 2); run<12>("I2F+IMAD", 2); run<13>("I2F+FFMA2", 2); run<14>("IMAD+FADD", 2); run<15>("LDS+FADD+STS", 3); run<16>("IMAD+I2F+FFMA+FADD+FMUL", 5);
  return 0;
}
