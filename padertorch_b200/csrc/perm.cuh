// Permutation search shared by the PIT kernels: all K! assignments of a K x K cost matrix in
// itertools.permutations order, first minimum wins (padertorch/ops/losses/source_separation.py:112-122).
#pragma once
#include "common.cuh"

namespace b2s {

__device__ __forceinline__ int factorial(int n) {
  int f = 1;
  for (int i = 2; i <= n; ++i) f *= i;
  return f;
}

// idx-th permutation of range(K) in lexicographic order (factorial number system).
__device__ __forceinline__ void unrank_permutation(int idx, int K, int* p) {
  int avail[B2S_MAX_SOURCES];
  for (int i = 0; i < K; ++i) avail[i] = i;
  for (int k = 0; k < K; ++k) {
    const int f = factorial(K - 1 - k);
    const int d = idx / f;
    idx -= d * f;
    p[k] = avail[d];
    for (int a = d; a < K - 1 - k; ++a) avail[a] = avail[a + 1];
  }
}

// torch.min semantics: NaN wins over numbers; among equals the lower index wins.
__device__ __forceinline__ bool candidate_better(double va, int ia, double vb, int ib) {
  const bool na = va != va, nb = vb != vb;
  if (na != nb) return na;
  if (!na && va != vb) return va < vb;
  return ia < ib;
}

// Called by ALL threads of a block (blockDim.x * blockDim.y * blockDim.z <= 1024).
// cost[i*K + j] (shared or global memory, already visible to the block) = cost of assigning estimate i
// to target j.  value(perm) = sum_k cost[perm[k]*K + k].  Thread 0 receives the winner.
__device__ inline void search_permutations(const double* cost, int K, double& best_value, int* best_perm) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int total = factorial(K);
  double val = 0.0;
  int idx = 0x7fffffff;
  bool have = false;
  for (int c = tid; c < total; c += nthreads) {
    int p[B2S_MAX_SOURCES];
    unrank_permutation(c, K, p);
    double v = 0.0;
    for (int k = 0; k < K; ++k) v += cost[p[k] * K + k];
    if (!have || candidate_better(v, c, val, idx)) { val = v; idx = c; have = true; }
  }
  // warp argmin; lanes without a candidate carry idx = INT_MAX and lose every comparison below
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, val, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    const bool ohave = oi != 0x7fffffff;
    if (ohave && (idx == 0x7fffffff || candidate_better(ov, oi, val, idx))) { val = ov; idx = oi; }
  }
  __syncthreads();
  if ((tid & 31) == 0) { s_val[tid >> 5] = val; s_idx[tid >> 5] = idx; }
  __syncthreads();
  if (tid == 0) {
    const int nwarps = (nthreads + 31) >> 5;
    val = s_val[0]; idx = s_idx[0];
    for (int w = 1; w < nwarps; ++w) {
      if (s_idx[w] != 0x7fffffff && (idx == 0x7fffffff || candidate_better(s_val[w], s_idx[w], val, idx))) {
        val = s_val[w]; idx = s_idx[w];
      }
    }
    best_value = val;
    unrank_permutation(idx, K, best_perm);
  }
}

}  // namespace b2s
