// Throughput of the prototype pair transform (tools/prototypes/cfft_pair.cuh: two real frames per 1024-point
// complex FFT, one exchange) with NO global traffic in the loop: frames resident in shared memory, 2 CTAs x 4
// warps per SM.  Compare with rfft_rate.cu (the transform in the tree in the same setting) and with the SM-side
// floor of the fused kernel (profiles/r1_fused_floor.txt: 43.6 us for 32 384 transforms + SSE = 1.35 ns per frame).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pair_rate pair_rate.cu && ./pair_rate
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../padertorch_b200/csrc/cfft_pair.cuh"
using namespace b2s::cp;

constexpr int kWarps = 4;
constexpr int kWarpFloats = 2048 + 2 * 32 * kPitch;

__global__ void __launch_bounds__(32 * kWarps, 2)
pair_rate_kernel(const float* __restrict__ x, const float2* __restrict__ tab, const float* __restrict__ win,
                 float* __restrict__ out, int iters) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* fa = sm + warp * kWarpFloats;
  float* fb = fa + 1024;
  float2* tile = reinterpret_cast<float2*>(fb + 1024);
  PairConsts k;
  k.init(tab, win, lane);
  for (int i = lane; i < 1024; i += 32) { fa[i] = x[i]; fb[i] = x[1024 + i]; }
  __syncwarp();
  float acc = 0.f;
  const int partner = (32 - lane) & 31;
  for (int it = 0; it < iters; ++it) {
    pass1(fa, fb, tile, k);
    __syncwarp();
    float2 z[32];
    pass2(tile, lane, z);
    __syncwarp();
    float2 m[16], sa[16], sb[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 send = lane == 0 ? z[(32 - r) & 31] : z[31 - r];
      m[r].x = __shfl_sync(0xffffffffu, send.x, partner);
      m[r].y = __shfl_sync(0xffffffffu, send.y, partner);
    }
    separate(z, m, sa, sb);
#pragma unroll
    for (int r = 0; r < 16; ++r) acc += sa[r].x * sa[r].x + sa[r].y * sa[r].y + sb[r].x * sb[r].x + sb[r].y * sb[r].y;
    if (lane == 0) fa[it & 1023] += 1e-9f * acc;   // keeps the loop alive, perturbs the input a little
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int sms = 148, grid = sms * 2, iters = 200;
  std::vector<float> hx(2048), hw(1024);
  std::vector<float2> htab(1024);
  for (int i = 0; i < 2048; ++i) hx[i] = (float)rand() / RAND_MAX - 0.5f;
  for (int i = 0; i < 1024; ++i) {
    hw[i] = (float)(0.42 - 0.5 * cos(2 * M_PI * i / 1024.0) + 0.08 * cos(4 * M_PI * i / 1024.0));
    htab[i] = make_float2((float)cos(-2 * M_PI * i / 1024.0), (float)sin(-2 * M_PI * i / 1024.0));
  }
  float *x, *w, *out; float2* tab;
  cudaMalloc(&x, 2048 * 4); cudaMalloc(&w, 1024 * 4); cudaMalloc(&tab, 1024 * 8); cudaMalloc(&out, grid * 32 * kWarps * 4);
  cudaMemcpy(x, hx.data(), 2048 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(w, hw.data(), 1024 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(tab, htab.data(), 1024 * 8, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(float) * kWarps * kWarpFloats;
  cudaFuncSetAttribute(pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_rate_kernel, 32 * kWarps, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    pair_rate_kernel<<<grid, 32 * kWarps, smem>>>(x, tab, w, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double frames = 2.0 * iters * grid * kWarps;
    printf("pair transform: %d CTAs/SM resident, %.1f us for %.0f frames -> %.3f ns per frame chip-wide (%s)\n", per_sm,
           ms * 1e3, frames, ms * 1e6 / frames, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
