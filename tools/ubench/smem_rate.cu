// Micro-benchmark: shared-memory instruction throughput per SM (LDS.32/64/128, STS.64/128, mixed), conflict free.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 256, UNROLL = 16;

template <int MODE>
__global__ void kern(float* out, long long* cycles) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* base = sm + warp * 2048;   // 8 KB per warp
  for (int i = lane; i < 2048; i += 32) base[i] = i;
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(base);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned a32 = sbase + 4 * (lane + 32 * u), a64 = sbase + 8 * (lane + 32 * u), a128 = sbase + 16 * (lane + 32 * u);
      float4 v;
      if (MODE == 0) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v.x) : "r"(a32)); acc.x += v.x; }
      if (MODE == 1) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a64)); acc.x += v.x; acc.y += v.y; }
      if (MODE == 2) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a128)); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      if (MODE == 3) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a64), "f"(acc.x), "f"(acc.y) : "memory"); }
      if (MODE == 4) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a128), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory"); }
      if (MODE == 5) {
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a128), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
        const unsigned r128 = sbase + 16 * (((lane + u) & 31) + 32 * u);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(r128));
        acc.x += v.x;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int bytes_per_lane, int instr_per_u) {
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  cudaFuncSetAttribute(kern<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int warps : {4, 8, 16}) {
    kern<MODE><<<148, 32 * warps, warps * 8192>>>(out, cyc);
    kern<MODE><<<148, 32 * warps, warps * 8192>>>(out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr = (double)ITERS * UNROLL * instr_per_u * warps;
    printf("%-22s warps/SM=%2d  cycles per warp-instr per SM = %6.3f   bytes/clk/SM = %6.1f\n", name, warps,
           avg / instr, instr * 32 * bytes_per_lane / avg);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  // (plain load loops are hoisted out of the timing loop by ptxas; loads are measured together with stores)
  run<3>("STS.64", 8, 1);
  run<4>("STS.128", 16, 1);
  run<5>("STS.128 + LDS.128", 16, 2);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
