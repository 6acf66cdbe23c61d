// Shared helpers of libb200sep (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200sep.h"

namespace b2s {

// thread-local error text returned by b2s_last_error()
void set_error(const char* fmt, ...);

#define B2S_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::b2s::set_error(__VA_ARGS__);      \
      return B2S_ERR_ARGUMENT;            \
    }                                     \
  } while (0)

#define B2S_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t err__ = (call);                                                     \
    if (err__ != cudaSuccess) {                                                     \
      ::b2s::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),   \
                       __FILE__, __LINE__);                                         \
      return B2S_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define B2S_LAUNCH_CHECK(name)                                                      \
  do {                                                                              \
    cudaError_t err__ = cudaGetLastError();                                         \
    if (err__ != cudaSuccess) {                                                     \
      ::b2s::set_error("launch of %s failed: %s", name, cudaGetErrorString(err__)); \
      return B2S_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: a C entry
// point must not leave a different current device behind (callers other than the Python wrappers rely on it).
struct DeviceGuard {
  int previous = -1;
  cudaError_t status;
  explicit DeviceGuard(int device) {
    status = cudaGetDevice(&previous);
    if (status == cudaSuccess && previous != device) status = cudaSetDevice(device); else if (status == cudaSuccess) previous = -1;
  }
  ~DeviceGuard() { if (previous >= 0) cudaSetDevice(previous); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define B2S_ON_DEVICE(device)                 \
  ::b2s::DeviceGuard device_guard__(device);  \
  B2S_CUDA(device_guard__.status)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Reduction workspaces: [kMaxTickets int ticket counters][double partial sums].  The counters sit at a
// FIXED place so that they stay zero across calls with different geometries (every kernel resets the
// counters it used); the partial-sum area may hold stale values, it is always written before it is read.
constexpr int64_t kMaxTickets = 1 << 20;
constexpr int64_t kTicketBytes = kMaxTickets * (int64_t)sizeof(int);
inline int* ws_counters(void* ws) { return reinterpret_cast<int*>(ws); }
inline double* ws_partials(void* ws) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kTicketBytes);
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide deterministic sum of NV values per thread; the result lands in out[0..NV) of thread 0's
// view via shared memory `red` (>= NV * warps doubles).  Fixed tree: lanes by xor-shuffle, warps in order.
template <int NV>
__device__ __forceinline__ void block_sum_to_smem(const float (&v)[NV], double* red, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = warp_sum(v[i]);
    if (lane == 0) red[i * nwarps + warp] = (double)s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += red[i * nwarps + w];
    out[i] = s;
  }
  __syncthreads();
}

// Next lexicographic permutation in place (the order of itertools.permutations(range(K))).
// p[0] + p[stride] + ... (n values) added in index order, four loads in flight (the values come from L2: one round
// trip per four values instead of one per value -- the fold of the chunk partials is a serial tail of every launch)
__device__ __forceinline__ double ordered_sum(const volatile double* p, int n, int64_t stride) {
  double s = 0.0;
  int c = 0;
  for (; c + 4 <= n; c += 4) {
    const double d0 = p[(int64_t)c * stride];
    const double d1 = p[(int64_t)(c + 1) * stride];
    const double d2 = p[(int64_t)(c + 2) * stride];
    const double d3 = p[(int64_t)(c + 3) * stride];
    s += d0; s += d1; s += d2; s += d3;
  }
  for (; c < n; ++c) s += p[(int64_t)c * stride];
  return s;
}

__device__ __forceinline__ bool next_permutation(int* p, int K) {
  int i = K - 2;
  while (i >= 0 && p[i] > p[i + 1]) --i;
  if (i < 0) return false;
  int j = K - 1;
  while (p[j] < p[i]) --j;
  int t = p[i]; p[i] = p[j]; p[j] = t;
  for (int a = i + 1, b = K - 1; a < b; ++a, --b) { t = p[a]; p[a] = p[b]; p[b] = t; }
  return true;
}

}  // namespace b2s
