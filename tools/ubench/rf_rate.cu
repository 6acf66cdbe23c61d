// Micro-benchmark: do packed instructions with all-distinct register operands (no operand reuse) run slower?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
constexpr int N = 24, ITERS = 512;
template <int MODE>
__global__ void kern(float* out, long long* cycles, float seed) {
  float2 v[N];
  for (int i = 0; i < N; ++i) v[i] = make_float2(seed + i + threadIdx.x, seed - i);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      // butterfly-like: every instruction reads two or three DIFFERENT live registers
      if (MODE == 0) v[i] = add2(v[(i + 5) % N], v[(i + 11) % N]);
      if (MODE == 1) v[i] = fma2(v[(i + 5) % N], v[(i + 11) % N], v[(i + 17) % N]);
      if (MODE == 2) { v[i].x = v[(i + 5) % N].x + v[(i + 11) % N].y; v[i].y = v[(i + 5) % N].y - v[(i + 11) % N].x; }
      if (MODE == 3) { v[i].x = fmaf(v[(i + 5) % N].x, v[(i + 11) % N].x, v[(i + 17) % N].y);
                       v[i].y = fmaf(v[(i + 5) % N].y, v[(i + 11) % N].y, v[(i + 17) % N].x); }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < N; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int per) {
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * 148 * 1024); cudaMalloc(&cyc, sizeof(long long) * 148);
  for (int w : {1, 2, 4}) {
    kern<MODE><<<148, 128 * w>>>(out, cyc, 1e-3f); kern<MODE><<<148, 128 * w>>>(out, cyc, 1e-3f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-40s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, w, avg / ((double)ITERS * N * per * w));
  }
}
int main() {
  run<0>("FADD2, two distinct source pairs", 1);
  run<1>("FFMA2, three distinct source pairs", 1);
  run<2>("2 x FADD scalar, distinct sources", 2);
  run<3>("2 x FFMA scalar, three distinct sources", 2);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
