// Throughput of the transform in the tree (rfft_packed.cuh: rfft_streams<2>, 8 x 8 x 8 with two exchanges and the
// real split) in the same setting as pair_rate.cu: two frames per call resident in shared memory, no global traffic
// in the loop, 2 CTAs x 4 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o rfft_rate rfft_rate.cu && ./rfft_rate
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../padertorch_b200/csrc/rfft_packed.cuh"
using namespace b2s;

constexpr int kWarps = 4;
constexpr int kWarpFloats = 2 * rf::kSize + 4 * rf::kTile1;

__global__ void __launch_bounds__(32 * kWarps, 3)
rfft_rate_kernel(const float* __restrict__ x, const float4* __restrict__ lane_table, float* __restrict__ out, int iters) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sig = sm + warp * kWarpFloats;
  float2* tile = reinterpret_cast<float2*>(sig + 2 * rf::kSize);
  rf::LaneConsts k;
  k.load(lane_table, lane);
  for (int i = lane; i < 2048; i += 32) sig[i] = x[i];
  __syncwarp();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float2 ya[2][8], yb[2][8];
    float ydc[2], ynyq[2];
    rf::rfft_streams<2, false, false>(sig, rf::kSize, tile, k, ya, yb, ydc, ynyq);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int p = 0; p < 8; ++p)
        acc += ya[s][p].x * ya[s][p].x + ya[s][p].y * ya[s][p].y + yb[s][p].x * yb[s][p].x + yb[s][p].y * yb[s][p].y;
    if (lane == 0) sig[it & 1023] += 1e-9f * (acc + ydc[0] + ynyq[1]);
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
  // argv[1] = resident CTAs per SM wanted (2 or 3: dynamic shared memory is padded so that no more fit)
  const int want = argc > 1 ? atoi(argv[1]) : 3;
  const int sms = 148, grid = sms * want, iters = 200;
  std::vector<float> hx(2048), hw(1024);
  std::vector<float2> htab(1024);
  for (int i = 0; i < 2048; ++i) hx[i] = (float)rand() / RAND_MAX - 0.5f;
  for (int i = 0; i < 1024; ++i) {
    hw[i] = (float)(0.42 - 0.5 * cos(2 * M_PI * i / 1024.0) + 0.08 * cos(4 * M_PI * i / 1024.0));
    htab[i] = make_float2((float)cos(-2 * M_PI * i / 1024.0), (float)sin(-2 * M_PI * i / 1024.0));
  }
  std::vector<float4> table(rf::kConstFloat4 * 32);
  for (int l = 0; l < 32; ++l) {
    rf::LaneConsts c;
    c.init(htab.data(), hw.data(), l);
    c.pack(table.data());
  }
  float *x, *out; float4* tab;
  cudaMalloc(&x, 2048 * 4); cudaMalloc(&tab, table.size() * 16); cudaMalloc(&out, grid * 32 * kWarps * 4);
  cudaMemcpy(x, hx.data(), 2048 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(tab, table.data(), table.size() * 16, cudaMemcpyHostToDevice);
  size_t smem = sizeof(float) * kWarps * kWarpFloats;
  if (want == 2) smem = 100 * 1024;   // two CTAs of 100 KB fit, three do not
  cudaFuncSetAttribute(rfft_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rfft_rate_kernel, 32 * kWarps, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    rfft_rate_kernel<<<grid, 32 * kWarps, smem>>>(x, tab, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double frames = 2.0 * iters * grid * kWarps;
    printf("transform in the tree: %d CTAs/SM resident, %.1f us for %.0f frames -> %.3f ns per frame chip-wide (%s)\n",
           per_sm, ms * 1e3, frames, ms * 1e6 / frames, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
