// Micro-benchmark: issue rate of packed fp32x2 versus scalar fp32 instructions on one SM sub-partition.
// Each warp runs a loop of N independent dependency chains; cycles per instruction per SMSP are printed for
// 1, 2 and 4 warps per SMSP.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rate pipe_rate.cu
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
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fadd(float a, float b) {
  float r;
  asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

constexpr int CH = 16, ITERS = 512;

template <int MODE>
__global__ void kern(float* out, long long* cycles, float seed) {
  float2 acc[CH];
  for (int i = 0; i < CH; ++i) acc[i] = make_float2(seed + i + threadIdx.x, seed - i);
  const float2 b = make_float2(seed * 0.5f, seed * 0.25f), c = make_float2(seed, -seed);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) acc[i] = fma2(acc[i], b, c);                                   // FFMA2
      if (MODE == 1) acc[i] = add2(acc[i], c);                                      // FADD2
      if (MODE == 2) acc[i] = add2(acc[i], make_float2(acc[(i + 1) % CH].y, -acc[(i + 1) % CH].x));  // FADD2 LO_HI.NP
      if (MODE == 3) { acc[i].x = ffma(acc[i].x, b.x, c.x); acc[i].y = ffma(acc[i].y, b.y, c.y); }   // 2 x FFMA
      if (MODE == 4) { acc[i].x = fadd(acc[i].x, c.x); acc[i].y = fadd(acc[i].y, c.y); }             // 2 x FADD
      if (MODE == 5) acc[i] = fma2(acc[i], make_float2(b.x, b.x), c);               // FFMA2 with .F32 broadcast
      if (MODE == 6) acc[i] = add2(acc[i], make_float2(acc[(i + 1) % CH].x, -acc[(i + 1) % CH].y));  // FADD2 HI_LO.NP (conj)
      if (MODE == 7) acc[i] = add2(acc[i], make_float2(-acc[(i + 1) % CH].x, -acc[(i + 1) % CH].y)); // FADD2 -R
      if (MODE == 8) acc[i] = add2(acc[i], make_float2(acc[(i + 1) % CH].y, acc[(i + 1) % CH].x));   // FADD2 LO_HI (swap only)
      if (MODE == 9) acc[i] = fma2(make_float2(acc[i].y, acc[i].y), b, c);          // FFMA2 broadcast of the high half
      if (MODE == 10) acc[i] = add2(acc[i], acc[(i + 1) % CH]);                     // FADD2 reg+reg (other chain)
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < CH; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_chain_step) {
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    kern<MODE><<<148, threads>>>(out, cyc, 1.0f);
    kern<MODE><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr_per_warp = (double)ITERS * CH * instr_per_chain_step;
    printf("%-28s warps/SMSP=%d  cycles=%9.0f  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, avg,
           avg / (instr_per_warp * warps_per_smsp));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA2 (16 chains)", 1);
  run<1>("FADD2", 1);
  run<2>("FADD2 LO_HI.NP operand", 1);
  run<3>("FFMA scalar (x2)", 2);
  run<4>("FADD scalar (x2)", 2);
  run<5>("FFMA2 .F32 broadcast", 1);
  run<6>("FADD2 HI_LO.NP (conj)", 1);
  run<7>("FADD2 -R (negate)", 1);
  run<8>("FADD2 LO_HI (swap only)", 1);
  run<9>("FFMA2 hi-half broadcast", 1);
  run<10>("FADD2 reg + other reg", 1);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
