// What does clock64() tick at?  Compares it with %globaltimer (ns) over a busy loop, and measures the issue rate
// of independent scalar FADDs / IADDs per SMSP in clock64 units.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out) {
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long c0 = clock64();
  float a = threadIdx.x, b = 1.0001f;
  for (int i = 0; i < 2000000; ++i) a = a * b + 0.5f;
  const long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(g1 - g0); out[2] = (long long)a; }
}
template <int MODE>
__global__ void rate(long long* out, int n) {
  int x[16]; float f[16];
  for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x + i; f[i] = threadIdx.x + i; }
  __syncthreads();
  const long long c0 = clock64();
#pragma unroll 1
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(it));
      if (MODE == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.5f));
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  int s = 0; for (int i = 0; i < 16; ++i) s += x[i] + (int)f[i];
  if (threadIdx.x == 0) { out[0] = c1 - c0; out[1] = s; }
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[3];
  k<<<1, 32>>>(d); cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("clock64 ticks %lld in %lld ns -> %.3f GHz\n", h[0], h[1], (double)h[0] / h[1]);
  for (int w : {1, 2, 4}) {
    rate<0><<<148, 128 * w>>>(d, 4096); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("IADD  warps/SMSP=%d: %.3f clock64 ticks per warp-instr per SMSP\n", w, (double)h[0] / (4096.0 * 16 * w));
    rate<1><<<148, 128 * w>>>(d, 4096); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("FADD  warps/SMSP=%d: %.3f clock64 ticks per warp-instr per SMSP\n", w, (double)h[0] / (4096.0 * 16 * w));
  }
  return 0;
}
