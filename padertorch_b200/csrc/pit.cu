// Permutation-invariant MSE over spectrogram-shaped blocks: fused mask (*) observation, K x K
// sum-of-squared-error matrix, permutation search -- one launch for a whole ragged batch.
// Reference: padertorch/ops/losses/source_separation.py:34-124 driven by the per-example loop of
// padertorch/contrib/examples/source_separation/pit/model.py:117-135 (which materialises mask*Y twice
// and launches ~20 ATen kernels per example).
//
// HBM-bound streaming reduction: every input element is read exactly once with coalesced 4-byte
// loads (F = 513 rows are not 16-byte aligned), two frames in flight per thread; partial sums leave
// the CTA as doubles and the last CTA of an example (ticket counter) folds them in a fixed order.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "perm.cuh"

using namespace b2s;

namespace {

constexpr int kPitWarps = 8;
constexpr int kPitThreads = 32 * kPitWarps;
constexpr int kPitUnroll = 4;          // bin groups of 32 in flight per lane
constexpr int kPitCtasPerSm = 3;

struct PitGrid {
  int tchunks, fchunks;
  int nchunks() const { return tchunks * fchunks; }
};

// Upstream gradient of the backward kernels: value (slot, example) = scale * p[slot * slot_stride + example * stride];
// stride 0 broadcasts one value per slot (the gradient of a batch mean, scale 1 / batch: b2s_pit_sse_backward_scaled).
struct GradLoss {
  const float* p; int64_t stride, slot_stride; double scale;
  __device__ __forceinline__ double at(int slot, int64_t b) const { return scale * (double)p[slot * slot_stride + b * stride]; }
};

constexpr int64_t kBinsPerChunk = 16384;

PitGrid pit_grid(int64_t batch, int64_t max_frames, int64_t bins) {
  PitGrid g;
  g.fchunks = (int)ceil_div(std::max<int64_t>(bins, 1), kBinsPerChunk);
  // when there are few frames but very long rows (pit_loss on [K, T] signals) split the rows instead
  if (max_frames < 4 * kPitWarps && bins > 4096)
    g.fchunks = (int)std::min<int64_t>(ceil_div(bins, 2048), 1024);
  const int64_t capacity = (int64_t)kNumSMs * kPitCtasPerSm;
  int64_t t = capacity / std::max<int64_t>(1, batch * g.fchunks);
  const int64_t most = std::max<int64_t>(1, max_frames / kPitWarps);
  g.tchunks = (int)std::max<int64_t>(1, std::min<int64_t>(t, most));
  static const int forced = [] { const char* e = getenv("B2S_PIT_TCHUNKS"); return e ? atoi(e) : 0; }();
  if (forced > 0) g.tchunks = (int)std::min<int64_t>(forced, most);   // tuning aid
  return g;
}

// CTA reduction (fixed tree) -> partial[b][chunk][NV]; the last CTA of an example (ticket) folds the chunks in
// order and picks the permutation.  `sm`: shared scratch of NV * (warps + 1) doubles.  All threads call it.
template <int K, bool DUAL>
__device__ __forceinline__ void pit_finish(const float (&acc)[(DUAL ? 2 : 1) * K * K], double* sm, int b, int chunk,
                                           int nchunks, int64_t T, int64_t F, double* __restrict__ partial,
                                           int* __restrict__ counters, float* __restrict__ loss,
                                           int32_t* __restrict__ perm, double* __restrict__ sse, int64_t batch) {
  constexpr int NV = (DUAL ? 2 : 1) * K * K;
  const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
  // ---- CTA reduction (fixed tree) -> partial[b][chunk][NV]
  const int tid = threadIdx.x;
  constexpr int nthreads = kPitThreads;
  const int lane = lane_, warp = warp_;
  constexpr int nwarps = kPitWarps;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) sm[i * nwarps + warp] = (double)s;
  }
  __syncthreads();
  double* mine = partial + ((int64_t)b * nchunks + chunk) * NV;
  for (int i = tid; i < NV; i += nthreads) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += sm[i * nwarps + w];
    mine[i] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (tid == 0) s_last = atomicAdd(counters + b, 1) == nchunks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- last CTA of this example: fold the chunks in order, pick the permutation
  double* total = sm + NV * nwarps;
  for (int i = tid; i < NV; i += nthreads) {
    const double s = ordered_sum(partial + (int64_t)b * nchunks * NV + i, nchunks, NV);
    total[i] = s;
    sse[(int64_t)b * NV + i] = s;
  }
  __syncthreads();
  const double count = (double)T * (double)K * (double)F;
#pragma unroll
  for (int slot = 0; slot < (DUAL ? 2 : 1); ++slot) {
    double best;
    int bp[B2S_MAX_SOURCES];
    search_permutations(total + slot * K * K, K, best, bp);
    if (tid == 0) {
      loss[slot * batch + b] = (float)(best / count);
      for (int k = 0; k < K; ++k) perm[(slot * batch + b) * K + k] = bp[k];
    }
    __syncthreads();
  }
  if (tid == 0) counters[b] = 0;  // leave the workspace ready for the next call
}

template <int K, bool DUAL>
__global__ void __launch_bounds__(kPitThreads, kPitCtasPerSm)
pit_sse_forward_kernel(const float* __restrict__ mask, const float* __restrict__ obs,
                       const float* __restrict__ tgt, const float* __restrict__ scale,
                       const int64_t* __restrict__ meta, int tchunks, int fchunks, int64_t F,
                       double* __restrict__ partial, int* __restrict__ counters,
                       float* __restrict__ loss, int32_t* __restrict__ perm, double* __restrict__ sse,
                       int64_t batch) {
  constexpr int NV = (DUAL ? 2 : 1) * K * K;
  extern __shared__ double sm[];  // [NV * nwarps] reduction scratch, then [NV] totals
  const int b = blockIdx.x;
  const int chunk = blockIdx.y, nchunks = tchunks * fchunks;
  const int tc = chunk / fchunks, fc = chunk - tc * fchunks;
  const int64_t T = meta[b * B2S_PIT_META + 0];
  const float* m_ = mask + meta[b * B2S_PIT_META + 1];
  const float* o_ = obs ? obs + meta[b * B2S_PIT_META + 2] : nullptr;
  const float* x_ = tgt + meta[b * B2S_PIT_META + 3];
  const float* s_ = scale ? scale + meta[b * B2S_PIT_META + 4] : nullptr;
  const int64_t t0 = T * tc / tchunks, t1 = T * (tc + 1) / tchunks;
  const int64_t f0 = F * fc / fchunks, f1 = F * (fc + 1) / fchunks;
  const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;

  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  // a warp owns frames t0 + warp, t0 + warp + 8, ...; its lanes sweep the bins, kPitUnroll groups of 32
  // bins (= kPitUnroll * (1 + 2K [+K]) independent coalesced loads) in flight
  for (int64_t t = t0 + warp_; t < t1; t += kPitWarps) {
    const float* mrow = m_ + t * K * F;
    const float* xrow = x_ + t * K * F;
    const float* srow = s_ ? s_ + t * K * F : nullptr;
    const float* orow = o_ ? o_ + t * F : nullptr;
    for (int64_t fb = f0 + lane_; fb < f1; fb += 32 * kPitUnroll) {
      float o[kPitUnroll], e[kPitUnroll][K], x[kPitUnroll][K], sc[kPitUnroll][K];
#pragma unroll
      for (int u = 0; u < kPitUnroll; ++u) {
        const int64_t f = fb + 32 * u;
        const bool on = f < f1;
        o[u] = (on && orow) ? __ldg(orow + f) : (on ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < K; ++i) {
          e[u][i] = on ? __ldg(mrow + i * F + f) : 0.f;
          x[u][i] = on ? __ldg(xrow + i * F + f) : 0.f;
          sc[u][i] = (on && srow) ? __ldg(srow + i * F + f) : 1.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kPitUnroll; ++u) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float ei = e[u][i] * o[u];
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if (DUAL) {
              const float d0 = ei - x[u][j], d1 = ei - x[u][j] * sc[u][j];
              acc[i * K + j] = fmaf(d0, d0, acc[i * K + j]);
              acc[K * K + i * K + j] = fmaf(d1, d1, acc[K * K + i * K + j]);
            } else {
              const float d = ei - x[u][j] * sc[u][j];
              acc[i * K + j] = fmaf(d, d, acc[i * K + j]);
            }
          }
        }
      }
    }
  }

  pit_finish<K, DUAL>(acc, sm, b, chunk, nchunks, T, F, partial, counters, loss, perm, sse, batch);
}

// ------------------------------------------------------------------------------------------- frame-staged forward
// Consecutive frames of an example are contiguous in every operand ([T][K][F] mask / target / scale, [T][F]
// observation): a stage of G frames arrives by one TMA bulk copy per operand (enclosing 16-byte aligned range,
// the data sits at the source's misalignment), double buffered per CTA, two CTAs per SM.  Warps read their
// frame's rows with conflict-free LDS.32 (lanes own bins).  Rows of 513 floats are only 4-byte aligned: the
// direct kernel above reads them with 4-byte loads and stalls on every batch of them.
constexpr int kStageBudget = 100 * 1024;   // dynamic shared memory per CTA (two stages)

__host__ __device__ inline int stage_area(int64_t floats) { return (int)((floats + 3 + 3) / 4 * 4); }
__host__ __device__ inline int stage_floats(int G, int K, int64_t F, bool has_obs, bool has_scale) {
  return stage_area(G * K * F) * (has_scale ? 3 : 2) + (has_obs ? stage_area(G * F) : 0);
}
// frames per stage: the largest power of two <= 8 whose two stages fit the budget (0: use the direct kernel)
inline int stage_frames(int K, int64_t F, bool has_obs, bool has_scale) {
  for (int G = kPitWarps; G >= 1; G >>= 1)
    if ((int64_t)stage_floats(G, K, F, has_obs, has_scale) * 2 * 4 <= kStageBudget) return G;
  return 0;
}

template <int K, bool DUAL>
__global__ void __launch_bounds__(kPitThreads, 2)
pit_sse_frame_kernel(const float* __restrict__ mask, const float* __restrict__ obs,
                     const float* __restrict__ tgt, const float* __restrict__ scale,
                     const int64_t* __restrict__ meta, int tchunks, int G, int F,
                     double* __restrict__ partial, int* __restrict__ counters,
                     float* __restrict__ loss, int32_t* __restrict__ perm, double* __restrict__ sse,
                     int64_t batch) {
  constexpr int NV = (DUAL ? 2 : 1) * K * K;
  extern __shared__ __align__(16) float stage_sm[];   // [2][mask | target | (scale) | (observation)]
  __shared__ __align__(8) uint64_t full[2];
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t T = meta[b * B2S_PIT_META + 0];
  const float* m_ = mask + meta[b * B2S_PIT_META + 1];
  const float* o_ = obs ? obs + meta[b * B2S_PIT_META + 2] : nullptr;
  const float* x_ = tgt + meta[b * B2S_PIT_META + 3];
  const float* s_ = scale ? scale + meta[b * B2S_PIT_META + 4] : nullptr;
  const int t0 = (int)(T * chunk / tchunks), t1 = (int)(T * (chunk + 1) / tchunks);
  const int area_k = stage_area((int64_t)G * K * F), area_o = o_ ? stage_area((int64_t)G * F) : 0;
  const int off_x = area_k, off_s = 2 * area_k, off_o = (s_ ? 3 : 2) * area_k;
  const int per_stage = off_o + area_o;
  if (threadIdx.x == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  // stage n covers frames [t0 + n G, min(t1, t0 + (n + 1) G))
  const int nstages = (t1 - t0 + G - 1) / G;
  auto span = [&](const float* base, int64_t floats, float* dst, uint64_t* bar, unsigned& bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base);
    bytes = (unsigned)(((a & 15) + (size_t)floats * 4 + 15) & ~(size_t)15);
    tma::bulk_g2s(dst, reinterpret_cast<const void*>(a & ~(uintptr_t)15), bytes, bar);
  };
  auto issue = [&](int n) {   // thread 0
    const int s = n & 1;
    const int t = t0 + n * G, g = min(G, t1 - t);
    float* buf = stage_sm + s * per_stage;
    const int64_t kf = (int64_t)K * F;
    // expected bytes first: the sizes follow from the addresses
    auto bytes_of = [&](const float* base, int64_t floats) {
      return (unsigned)(((reinterpret_cast<uintptr_t>(base) & 15) + (size_t)floats * 4 + 15) & ~(size_t)15);
    };
    unsigned total = bytes_of(m_ + t * kf, g * kf) + bytes_of(x_ + t * kf, g * kf);
    if (s_) total += bytes_of(s_ + t * kf, g * kf);
    if (o_) total += bytes_of(o_ + (int64_t)t * F, (int64_t)g * F);
    tma::fence_proxy_async();   // the buffer was last read through the generic proxy
    tma::mbar_expect_tx(&full[s], total);
    unsigned nb;
    span(m_ + t * kf, g * kf, buf, &full[s], nb);
    span(x_ + t * kf, g * kf, buf + off_x, &full[s], nb);
    if (s_) span(s_ + t * kf, g * kf, buf + off_s, &full[s], nb);
    if (o_) span(o_ + (int64_t)t * F, (int64_t)g * F, buf + off_o, &full[s], nb);
  };
  if (threadIdx.x == 0) {
    if (nstages > 0) issue(0);
    if (nstages > 1) issue(1);
  }
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  // warp -> (frame of the stage, part of the bins): G frames x (8 / G) parts
  const int fi = warp % G, part = warp / G, parts = kPitWarps / G;
  const int steps = (F + 31) / 32;
  for (int n = 0; n < nstages; ++n) {
    const int s = n & 1;
    const int t = t0 + n * G, g = min(G, t1 - t);
    tma::mbar_wait(&full[s], (unsigned)(n >> 1) & 1u);
    if (fi < g) {
      const float* buf = stage_sm + s * per_stage;
      const int64_t kf = (int64_t)K * F;
      auto mis = [&](const float* base) { return (int)((reinterpret_cast<uintptr_t>(base) & 15) >> 2); };
      const float* mrow = buf + mis(m_ + t * kf) + fi * K * F + lane;
      const float* xrow = buf + off_x + mis(x_ + t * kf) + fi * K * F + lane;
      const float* srow = s_ ? buf + off_s + mis(s_ + t * kf) + fi * K * F + lane : nullptr;
      const float* orow = o_ ? buf + off_o + mis(o_ + (int64_t)t * F) + fi * F + lane : nullptr;
#pragma unroll 2
      for (int j = part; j < steps; j += parts) {
        if (lane + 32 * j < F) {
          const float o = orow ? orow[32 * j] : 1.f;
          float e[K], x[K], sc[K];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            e[i] = mrow[i * F + 32 * j] * o;
            x[i] = xrow[i * F + 32 * j];
            sc[i] = srow ? srow[i * F + 32 * j] : 1.f;
          }
#pragma unroll
          for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
              if (DUAL) {
                const float d0 = e[i] - x[jj], d1 = e[i] - x[jj] * sc[jj];
                acc[i * K + jj] = fmaf(d0, d0, acc[i * K + jj]);
                acc[K * K + i * K + jj] = fmaf(d1, d1, acc[K * K + i * K + jj]);
              } else {
                const float d = e[i] - x[jj] * sc[jj];
                acc[i * K + jj] = fmaf(d, d, acc[i * K + jj]);
              }
            }
          }
        }
      }
    }
    __syncthreads();   // every warp is done with buffer s
    if (threadIdx.x == 0 && n + 2 < nstages) issue(n + 2);
  }
  // the stage buffers are free (every copy was waited for): reuse them as the reduction scratch
  pit_finish<K, DUAL>(acc, reinterpret_cast<double*>(stage_sm), b, chunk, tchunks, T, (int64_t)F, partial, counters,
                      loss, perm, sse, batch);
}

// grad_mask[t,i,f] = sum_slot c_slot * o * (m*o - x_slot[inv_slot[i]]),  c = grad_loss * 2 / (T K F)
template <int K, bool DUAL>
__global__ void __launch_bounds__(kPitThreads, kPitCtasPerSm)
pit_sse_backward_kernel(const float* __restrict__ mask, const float* __restrict__ obs,
                        const float* __restrict__ tgt, const float* __restrict__ scale,
                        const int64_t* __restrict__ meta, int tchunks, int fchunks, int64_t F,
                        const int32_t* __restrict__ perm, const GradLoss grad_loss,
                        float* __restrict__ grad_mask, float* __restrict__ grad_target, int64_t batch) {
  const int b = blockIdx.x;
  const int chunk = blockIdx.y;
  const int tc = chunk / fchunks, fc = chunk - tc * fchunks;
  const int64_t T = meta[b * B2S_PIT_META + 0];
  const float* m_ = mask + meta[b * B2S_PIT_META + 1];
  const float* o_ = obs ? obs + meta[b * B2S_PIT_META + 2] : nullptr;
  const float* x_ = tgt + meta[b * B2S_PIT_META + 3];
  const float* s_ = scale ? scale + meta[b * B2S_PIT_META + 4] : nullptr;
  float* gm_ = grad_mask + meta[b * B2S_PIT_META + 5];
  float* gt_ = grad_target ? grad_target + meta[b * B2S_PIT_META + 5] : nullptr;
  const int64_t t0 = T * tc / tchunks, t1 = T * (tc + 1) / tchunks;
  const int64_t f0 = F * fc / fchunks, f1 = F * (fc + 1) / fchunks;
  const double count = (double)T * (double)K * (double)F;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inv0[K], inv1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i0 = perm[(int64_t)b * K + k];
#pragma unroll
    for (int i = 0; i < K; ++i) if (i == i0) inv0[i] = k;
    if (DUAL) {
      const int i1 = perm[(batch + b) * K + k];
#pragma unroll
      for (int i = 0; i < K; ++i) if (i == i1) inv1[i] = k;
    }
  }
  const float c0 = (float)(2.0 * grad_loss.at(0, b) / count);
  const float c1 = DUAL ? (float)(2.0 * grad_loss.at(1, b) / count) : 0.f;
  constexpr int U = 2;
  for (int64_t t = t0 + warp; t < t1; t += kPitWarps) {
    const float* mrow = m_ + t * K * F;
    const float* xrow = x_ + t * K * F;
    const float* srow = s_ ? s_ + t * K * F : nullptr;
    const float* orow = o_ ? o_ + t * F : nullptr;
    float* grow = gm_ + t * K * F;
    float* gtrow = gt_ ? gt_ + t * K * F : nullptr;
    for (int64_t fb = f0 + lane; fb < f1; fb += 32 * U) {
      float o[U], e[U][K], x[U][K], sc[U][K];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t f = fb + 32 * u;
        const bool on = f < f1;
        o[u] = (on && orow) ? __ldg(orow + f) : 1.f;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          e[u][i] = on ? __ldg(mrow + i * F + f) * o[u] : 0.f;
          x[u][i] = on ? __ldg(xrow + i * F + f) : 0.f;
          sc[u][i] = (on && srow) ? __ldg(srow + i * F + f) : 1.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t f = fb + 32 * u;
        if (f >= f1) continue;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          float xa = 0.f, xb = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            if (inv0[i] == k) xa = DUAL ? x[u][k] : x[u][k] * sc[u][k];
            if (DUAL && inv1[i] == k) xb = x[u][k] * sc[u][k];
          }
          float g = c0 * (e[u][i] - xa);
          if (DUAL) g = fmaf(c1, e[u][i] - xb, g);
          grow[i * F + f] = g * o[u];
        }
        if (gtrow) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            float ek = 0.f;
#pragma unroll
            for (int i = 0; i < K; ++i) if (inv0[i] == k) ek = e[u][i];
            gtrow[k * F + f] = -c0 * (ek - x[u][k] * sc[u][k]) * sc[u][k];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- frame-staged backward
// Inputs staged exactly as in pit_sse_frame_kernel; the stage's [G][K][F] gradient block -- contiguous in global
// memory -- is assembled in shared memory at the destination's phase within 16 bytes and leaves as ONE TMA bulk
// store (double buffered; at most 3 floats at either end from lanes).  grad_target is not produced here.
inline int stage_frames_bwd(int K, int64_t F, bool has_obs, bool has_scale) {
  for (int G = kPitWarps; G >= 1; G >>= 1)
    if (((int64_t)stage_floats(G, K, F, has_obs, has_scale) + stage_area((int64_t)G * K * F)) * 2 * 4 <= kStageBudget)
      return G;
  return 0;
}

template <int K, bool DUAL>
__global__ void __launch_bounds__(kPitThreads, 2)
pit_sse_backward_frame_kernel(const float* __restrict__ mask, const float* __restrict__ obs,
                              const float* __restrict__ tgt, const float* __restrict__ scale,
                              const int64_t* __restrict__ meta, int tchunks, int G, int F,
                              const int32_t* __restrict__ perm, const GradLoss grad_loss,
                              float* __restrict__ grad_mask, int64_t batch) {
  extern __shared__ __align__(16) float stage_sm[];   // [2][mask | target | (scale) | (observation)], [2][gradient]
  __shared__ __align__(8) uint64_t full[2];
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t T = meta[b * B2S_PIT_META + 0];
  const float* m_ = mask + meta[b * B2S_PIT_META + 1];
  const float* o_ = obs ? obs + meta[b * B2S_PIT_META + 2] : nullptr;
  const float* x_ = tgt + meta[b * B2S_PIT_META + 3];
  const float* s_ = scale ? scale + meta[b * B2S_PIT_META + 4] : nullptr;
  float* gm_ = grad_mask + meta[b * B2S_PIT_META + 5];
  const int t0 = (int)(T * chunk / tchunks), t1 = (int)(T * (chunk + 1) / tchunks);
  const int area_k = stage_area((int64_t)G * K * F), area_o = o_ ? stage_area((int64_t)G * F) : 0;
  const int off_x = area_k, off_s = 2 * area_k, off_o = (s_ ? 3 : 2) * area_k;
  const int per_stage = off_o + area_o;
  float* outs = stage_sm + 2 * per_stage;
  const double count = (double)T * (double)K * (double)F;
  int inv0[K], inv1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i0 = perm[(int64_t)b * K + k];
#pragma unroll
    for (int i = 0; i < K; ++i) if (i == i0) inv0[i] = k;
    if (DUAL) {
      const int i1 = perm[(batch + b) * K + k];
#pragma unroll
      for (int i = 0; i < K; ++i) if (i == i1) inv1[i] = k;
    }
  }
  const float c0 = (float)(2.0 * grad_loss.at(0, b) / count);
  const float c1 = DUAL ? (float)(2.0 * grad_loss.at(1, b) / count) : 0.f;
  if (threadIdx.x == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  const int nstages = (t1 - t0 + G - 1) / G;
  const int64_t kf = (int64_t)K * F;
  auto issue = [&](int n) {   // thread 0
    const int s = n & 1;
    const int t = t0 + n * G, g = min(G, t1 - t);
    float* buf = stage_sm + s * per_stage;
    auto bytes_of = [&](const float* base, int64_t floats) {
      return (unsigned)(((reinterpret_cast<uintptr_t>(base) & 15) + (size_t)floats * 4 + 15) & ~(size_t)15);
    };
    auto span = [&](const float* base, int64_t floats, float* dst) {
      tma::bulk_g2s(dst, reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(base) & ~(uintptr_t)15),
                    bytes_of(base, floats), &full[s]);
    };
    unsigned total = bytes_of(m_ + t * kf, g * kf) + bytes_of(x_ + t * kf, g * kf);
    if (s_) total += bytes_of(s_ + t * kf, g * kf);
    if (o_) total += bytes_of(o_ + (int64_t)t * F, (int64_t)g * F);
    tma::fence_proxy_async();
    tma::mbar_expect_tx(&full[s], total);
    span(m_ + t * kf, g * kf, buf);
    span(x_ + t * kf, g * kf, buf + off_x);
    if (s_) span(s_ + t * kf, g * kf, buf + off_s);
    if (o_) span(o_ + (int64_t)t * F, (int64_t)g * F, buf + off_o);
  };
  if (threadIdx.x == 0) {
    if (nstages > 0) issue(0);
    if (nstages > 1) issue(1);
  }
  const int fi = warp % G, part = warp / G, parts = kPitWarps / G;
  const int steps = (F + 31) / 32;
  for (int n = 0; n < nstages; ++n) {
    const int s = n & 1;
    const int t = t0 + n * G, g = min(G, t1 - t);
    float* gout = gm_ + t * kf;                      // the stage's gradient block in global memory
    const int phase = (int)((reinterpret_cast<uintptr_t>(gout) & 15) >> 2);
    float* stage = outs + s * area_k + phase;        // stage[i] <-> gout[i]
    tma::mbar_wait(&full[s], (unsigned)(n >> 1) & 1u);
    if (fi < g) {
      const float* buf = stage_sm + s * per_stage;
      auto mis = [&](const float* base) { return (int)((reinterpret_cast<uintptr_t>(base) & 15) >> 2); };
      const float* mrow = buf + mis(m_ + t * kf) + fi * K * F + lane;
      const float* xrow = buf + off_x + mis(x_ + t * kf) + fi * K * F + lane;
      const float* srow = s_ ? buf + off_s + mis(s_ + t * kf) + fi * K * F + lane : nullptr;
      const float* orow = o_ ? buf + off_o + mis(o_ + (int64_t)t * F) + fi * F + lane : nullptr;
      float* grow = stage + fi * K * F + lane;
#pragma unroll 2
      for (int j = part; j < steps; j += parts) {
        if (lane + 32 * j < F) {
          const float o = orow ? orow[32 * j] : 1.f;
          float e[K], x[K], sc[K];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            e[i] = mrow[i * F + 32 * j] * o;
            x[i] = xrow[i * F + 32 * j];
            sc[i] = srow ? srow[i * F + 32 * j] : 1.f;
          }
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float xa = 0.f, xb = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              if (inv0[i] == k) xa = DUAL ? x[k] : x[k] * sc[k];
              if (DUAL && inv1[i] == k) xb = x[k] * sc[k];
            }
            float gv = c0 * (e[i] - xa);
            if (DUAL) gv = fmaf(c1, e[i] - xb, gv);
            grow[i * F + 32 * j] = gv * o;
          }
        }
      }
    }
    tma::fence_proxy_async();                          // the block was written through the generic proxy
    if (threadIdx.x == 0) tma::bulk_wait_read<0>();    // earlier stores have read their staging buffers
    __syncthreads();                                   // inputs of buffer s consumed, gradient block s complete
    const int nflt = (int)(g * kf);
    const int head = min(nflt, (4 - phase) & 3), mid = (nflt - head) & ~3, tail = nflt - head - mid;
    if (threadIdx.x == 0) {
      if (n + 2 < nstages) issue(n + 2);
      if (mid > 0) {
        tma::bulk_s2g(gout + head, stage + head, (unsigned)mid * 4u);
        tma::bulk_commit();
      }
    }
    if (warp == 1) {
      if (lane < head) gout[lane] = stage[lane];
      if (lane < tail) gout[head + mid + lane] = stage[head + mid + lane];
    }
  }
  if (threadIdx.x == 0) tma::bulk_wait<0>();   // shared memory must outlive the last store
}

template <int K>
int launch_forward_k(const float* mask, const float* obs, const float* tgt, const float* scale,
                     const int64_t* meta, int64_t batch, int64_t max_frames, int64_t F, int dual,
                     float* loss, int32_t* perm, double* sse, void* workspace, cudaStream_t stream) {
  const PitGrid g = pit_grid(batch, max_frames, F);
  const int nv = (dual ? 2 : 1) * K * K;
  (void)nv;
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  // frame-staged TMA path: rows short enough for whole frames to be staged, bins not split over CTAs
  static const bool no_frame = getenv("B2S_PIT_NO_FRAME") != nullptr;
  const int G = (g.fchunks == 1 && F < (1 << 20) && max_frames < (1 << 30))
      ? stage_frames(K, F, obs != nullptr, scale != nullptr) : 0;
  if (G > 0 && !no_frame && max_frames >= 2 * G) {
    const size_t smem = std::max<size_t>(sizeof(float) * 2 * stage_floats(G, K, F, obs != nullptr, scale != nullptr),
                                         sizeof(double) * nv * (kPitWarps + 1));
    auto kernel = dual ? pit_sse_frame_kernel<K, true> : pit_sse_frame_kernel<K, false>;
    static bool configured[2][64] = {};
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    if (!configured[dual ? 1 : 0][dev & 63]) {
      B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBudget));
      configured[dual ? 1 : 0][dev & 63] = true;
    }
    // the workspace holds g.nchunks() partial matrices per example: at most that many frame chunks
    const int tchunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(g.nchunks(), max_frames / (2 * G)),
                                                                   (int64_t)kNumSMs * 2 / std::max<int64_t>(1, batch)));
    kernel<<<dim3((unsigned)batch, tchunks), kPitThreads, smem, stream>>>(mask, obs, tgt, scale, meta, tchunks, G,
        (int)F, partial, counters, loss, perm, sse, batch);
    B2S_LAUNCH_CHECK("pit_sse_frame_kernel");
    return B2S_OK;
  }
  const dim3 grid((unsigned)batch, g.nchunks()), block(kPitThreads);
  const size_t smem = sizeof(double) * nv * (kPitWarps + 1);
  if (dual)
    pit_sse_forward_kernel<K, true><<<grid, block, smem, stream>>>(mask, obs, tgt, scale, meta, g.tchunks,
        g.fchunks, F, partial, counters, loss, perm, sse, batch);
  else
    pit_sse_forward_kernel<K, false><<<grid, block, smem, stream>>>(mask, obs, tgt, scale, meta, g.tchunks,
        g.fchunks, F, partial, counters, loss, perm, sse, batch);
  B2S_LAUNCH_CHECK("pit_sse_forward_kernel");
  return B2S_OK;
}

template <int K>
int launch_backward_k(const float* mask, const float* obs, const float* tgt, const float* scale,
                      const int64_t* meta, int64_t batch, int64_t max_frames, int64_t F, int dual,
                      const int32_t* perm, const GradLoss grad_loss, float* grad_mask,
                      float* grad_target, cudaStream_t stream) {
  static const bool no_frame = getenv("B2S_PIT_NO_FRAME") != nullptr;
  const int G = (!grad_target && F <= 16384 && max_frames < (1 << 30))
      ? stage_frames_bwd(K, F, obs != nullptr, scale != nullptr) : 0;
  if (G > 0 && !no_frame && max_frames >= 2 * G) {
    const size_t smem = sizeof(float) * 2 * ((size_t)stage_floats(G, K, F, obs != nullptr, scale != nullptr) +
                                             stage_area((int64_t)G * K * F));
    auto kernel = dual ? pit_sse_backward_frame_kernel<K, true> : pit_sse_backward_frame_kernel<K, false>;
    static bool configured[2][64] = {};
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    if (!configured[dual ? 1 : 0][dev & 63]) {
      B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBudget));
      configured[dual ? 1 : 0][dev & 63] = true;
    }
    const int tchunks = (int)std::max<int64_t>(1, std::min<int64_t>(max_frames / (2 * G),
                                                                   (int64_t)kNumSMs * 2 / std::max<int64_t>(1, batch)));
    kernel<<<dim3((unsigned)batch, tchunks), kPitThreads, smem, stream>>>(mask, obs, tgt, scale, meta, tchunks, G,
        (int)F, perm, grad_loss, grad_mask, batch);
    B2S_LAUNCH_CHECK("pit_sse_backward_frame_kernel");
    return B2S_OK;
  }
  PitGrid g = pit_grid(batch, max_frames, F);
  // elementwise: more CTAs than the reduction wants are fine
  g.tchunks = (int)std::max<int64_t>(g.tchunks, std::min<int64_t>(max_frames / (2 * kPitWarps) + 1, 64));
  const dim3 grid((unsigned)batch, g.nchunks()), block(kPitThreads);
  if (dual)
    pit_sse_backward_kernel<K, true><<<grid, block, 0, stream>>>(mask, obs, tgt, scale, meta, g.tchunks,
        g.fchunks, F, perm, grad_loss, grad_mask, grad_target, batch);
  else
    pit_sse_backward_kernel<K, false><<<grid, block, 0, stream>>>(mask, obs, tgt, scale, meta, g.tchunks,
        g.fchunks, F, perm, grad_loss, grad_mask, grad_target, batch);
  B2S_LAUNCH_CHECK("pit_sse_backward_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int64_t b2s_pit_workspace_bytes(int64_t batch, int64_t max_frames, int64_t bins, int sources, int dual) {
  if (batch <= 0 || sources <= 0 || bins <= 0) return 16;
  const PitGrid g = pit_grid(batch, max_frames, bins);
  const int64_t nv = (dual ? 2 : 1) * (int64_t)sources * sources;
  return kTicketBytes + (int64_t)sizeof(double) * batch * g.nchunks() * nv + 16;
}

int b2s_pit_sse_forward(const float* mask, const float* observation, const float* target,
                        const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                        int sources, int64_t bins, int dual, float* loss, int32_t* perm, double* sse,
                        void* workspace, b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES,
              "sources=%d outside the supported range 1..%d", sources, B2S_MAX_SOURCES);
  B2S_REQUIRE(batch >= 0 && batch <= kMaxTickets && bins >= 1 && max_frames >= 0, "bad extents");
  B2S_REQUIRE(!dual || scale, "dual PIT needs the target scale (cos phase difference)");
  if (batch == 0) return B2S_OK;
  B2S_REQUIRE(mask && target && meta && loss && perm && sse && workspace, "NULL device pointer");
#define CALL_FWD(K) launch_forward_k<K>(mask, observation, target, scale, meta, batch, max_frames, bins, \
                                        dual, loss, perm, sse, workspace, (cudaStream_t)stream)
  switch (sources) {
    case 1: return CALL_FWD(1);
    case 2: return CALL_FWD(2);
    case 3: return CALL_FWD(3);
    case 4: return CALL_FWD(4);
    case 5: return CALL_FWD(5);
    case 6: return CALL_FWD(6);
    case 7: return CALL_FWD(7);
    default: return CALL_FWD(8);
  }
#undef CALL_FWD
}

}  // extern "C"

namespace {
int pit_sse_backward_impl(const float* mask, const float* observation, const float* target,
                          const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                          int sources, int64_t bins, int dual, const int32_t* perm,
                          const GradLoss grad_loss, float* grad_mask, float* grad_target,
                          b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES,
              "sources=%d outside the supported range 1..%d", sources, B2S_MAX_SOURCES);
  B2S_REQUIRE(batch >= 0 && batch <= kMaxTickets && bins >= 1 && max_frames >= 0, "bad extents");
  B2S_REQUIRE(!(dual && grad_target), "grad_target is not available in dual mode");
  B2S_REQUIRE(!dual || scale, "dual PIT needs the target scale");
  if (batch == 0) return B2S_OK;
  B2S_REQUIRE(mask && target && meta && perm && grad_loss.p && grad_mask, "NULL device pointer");
#define CALL_BWD(K) launch_backward_k<K>(mask, observation, target, scale, meta, batch, max_frames, bins, \
                                         dual, perm, grad_loss, grad_mask, grad_target, (cudaStream_t)stream)
  switch (sources) {
    case 1: return CALL_BWD(1);
    case 2: return CALL_BWD(2);
    case 3: return CALL_BWD(3);
    case 4: return CALL_BWD(4);
    case 5: return CALL_BWD(5);
    case 6: return CALL_BWD(6);
    case 7: return CALL_BWD(7);
    default: return CALL_BWD(8);
  }
#undef CALL_BWD
}

// mean[slot] = mean_b loss[slot][b], fixed order (lane-strided partial sums, warp tree); one warp per slot
__global__ void __launch_bounds__(32)
pit_mean_kernel(const float* __restrict__ loss, int64_t batch, float* __restrict__ mean) {
  const float* row = loss + blockIdx.x * batch;
  double local = 0.0;
  for (int64_t i = threadIdx.x; i < batch; i += 32) local += (double)row[i];
  local = warp_sum(local);
  if (threadIdx.x == 0) mean[blockIdx.x] = (float)(local / (double)batch);
}
}  // namespace

extern "C" {

int b2s_pit_sse_backward(const float* mask, const float* observation, const float* target,
                         const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                         int sources, int64_t bins, int dual, const int32_t* perm,
                         const float* grad_loss, float* grad_mask, float* grad_target,
                         b2s_stream stream) {
  return pit_sse_backward_impl(mask, observation, target, scale, meta, batch, max_frames, sources, bins, dual, perm,
                               GradLoss{grad_loss, 1, batch, 1.0}, grad_mask, grad_target, stream);
}

int b2s_pit_sse_backward_scaled(const float* mask, const float* observation, const float* target,
                                const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                                int sources, int64_t bins, int dual, const int32_t* perm,
                                const float* grad_loss, int64_t grad_loss_stride, double grad_scale,
                                float* grad_mask, float* grad_target, b2s_stream stream) {
  B2S_REQUIRE(grad_loss_stride == 0 || grad_loss_stride == 1, "grad_loss_stride must be 0 (one value per slot) or 1");
  return pit_sse_backward_impl(mask, observation, target, scale, meta, batch, max_frames, sources, bins, dual, perm,
                               GradLoss{grad_loss, grad_loss_stride, grad_loss_stride ? batch : 1, grad_scale},
                               grad_mask, grad_target, stream);
}

int b2s_pit_sse_forward_mean(const float* mask, const float* observation, const float* target,
                             const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                             int sources, int64_t bins, int dual, float* loss, float* mean, int32_t* perm,
                             double* sse, void* workspace, b2s_stream stream) {
  B2S_REQUIRE(mean != nullptr && batch >= 1, "b2s_pit_sse_forward_mean needs batch >= 1 and a mean pointer");
  const int rc = b2s_pit_sse_forward(mask, observation, target, scale, meta, batch, max_frames, sources, bins, dual,
                                     loss, perm, sse, workspace, stream);
  if (rc != B2S_OK) return rc;
  pit_mean_kernel<<<dual ? 2 : 1, 32, 0, (cudaStream_t)stream>>>(loss, batch, mean);
  B2S_LAUNCH_CHECK("pit_mean_kernel");
  return B2S_OK;
}

}  // extern "C"
