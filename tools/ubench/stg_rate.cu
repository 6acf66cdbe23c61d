// Micro-benchmark: cost of coalesced global stores per SM (cycles per warp instruction), aligned vs misaligned
// 128-byte spans, 32/64/128-bit, with every SM streaming to its own region (write-back to L2 / HBM).
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH, int MISALIGN>   // WIDTH floats per lane, MISALIGN floats of offset
__global__ void kern(float* out, long long* cycles, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // each warp instruction stores 32 * WIDTH consecutive floats; consecutive instructions advance by 513 floats
  // when MISALIGN (like spectrum rows) or by 512 floats otherwise
  const long long pitch = MISALIGN ? 513 : 512;
  float* base = out + ((long long)blockIdx.x * nw + warp) * (long long)iters * 16 * pitch + MISALIGN;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      float* p = base + ((long long)it * 16 + u) * pitch + lane * WIDTH;
      if (WIDTH == 1) asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
      if (WIDTH == 2) asm volatile("st.global.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
      if (WIDTH == 4) asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int WIDTH, int MISALIGN>
void run(const char* name, float* buf, long long* cyc) {
  for (int warps : {4, 8, 16}) {
    const int iters = 64;
    kern<WIDTH, MISALIGN><<<148, 32 * warps>>>(buf, cyc, iters);
    cudaDeviceSynchronize();
    kern<WIDTH, MISALIGN><<<148, 32 * warps>>>(buf, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr = (double)iters * 16 * warps;
    printf("%-28s warps/SM=%2d  cycles per warp-store per SM = %6.2f  (%5.1f B/clk/SM, %6.0f GB/s chip)\n", name, warps,
           avg / instr, instr * 128 * WIDTH / avg, instr * 128 * WIDTH / avg * 148 * 1.965);
  }
}

int main() {
  float* buf; long long* cyc;
  const size_t bytes = (size_t)148 * 16 * 64 * 16 * 513 * 4 + 4096;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  run<1, 0>("STG.32 aligned (128 B)", buf, cyc);
  run<1, 1>("STG.32 misaligned +4 B", buf, cyc);
  run<2, 0>("STG.64 aligned (256 B)", buf, cyc);
  run<4, 0>("STG.128 aligned (512 B)", buf, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
