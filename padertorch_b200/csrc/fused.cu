// The fused north-star kernel: STFT -> |.| -> mask (*) |Y| -> K x K SSE -> permutation search, with the
// target spectra |STFT(s_k)| recomputed in registers instead of being written to and re-read from HBM.
// Equals  pit_loss(mask * Y_abs[:, None, :], X_abs, axis=-2)  with  X_abs = |STFT(s)|
// (padertorch/contrib/examples/source_separation/pit/data.py:49-77, pit/model.py:117-128,
//  padertorch/ops/losses/source_separation.py:34-124).
// Algorithmic HBM bytes per utterance: 4T(1+K) + 4MFK; reading the already
// materialised |Y| instead of recomputing it trades 4T for 4MF bytes against one FFT per frame.
//
// Work decomposition: a "group" is 4 consecutive frames of one example (one frame per warp of a 4-warp
// CTA).  The groups of the whole batch form one list; CTA c of a persistent grid owns a contiguous range
// of it, so every SM gets the same number of frames (+-1 group) and a CTA meets at most a few example
// boundaries, where it flushes its K x K partial sums (fixed-order reduction, ticket per example).
// The source samples of a group are staged ONCE in shared memory by zero-filling 16-byte cp.async
// (frames overlap 4x), double buffered against the transform of the previous group; mask and |Y| loads
// are issued before the first transform and consumed after it.
#include <algorithm>

#include "common.cuh"
#include "fft1024.cuh"
#include "stft_plan.cuh"
#include "perm.cuh"

using namespace b2s;

namespace {

constexpr int kFusedWarps = 4;
#ifndef B2S_FUSED_CTAS_PER_SM
#define B2S_FUSED_CTAS_PER_SM 3
#endif
constexpr int kFusedCtasPerSm = B2S_FUSED_CTAS_PER_SM;

struct FusedGrid {
  int grid;            // persistent CTAs
  int64_t gpe;         // groups per example (dense enumeration over `frames`)
  int64_t total;       // batch * gpe
  int slots;           // partial-sum slots per example
};

FusedGrid fused_grid(int64_t batch, int64_t frames) {
  FusedGrid g;
  g.gpe = std::max<int64_t>(1, ceil_div(frames, kFusedWarps));
  g.total = batch * g.gpe;
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>(g.total, (int64_t)kNumSMs * kFusedCtasPerSm));
  g.slots = (int)(ceil_div(g.gpe * g.grid, std::max<int64_t>(1, g.total)) + 2);
  return g;
}

// first group of CTA c: floor(c * total / grid); CTA owning group x: ceil((x + 1) * grid / total) - 1
__device__ __forceinline__ int64_t range_start(int64_t c, int64_t total, int64_t grid) {
  return c * total / grid;
}
__device__ __forceinline__ int64_t owner_of(int64_t x, int64_t total, int64_t grid) {
  return ((x + 1) * grid + total - 1) / total - 1;
}

template <int K, bool VEC16, bool MASK8>
__global__ void __launch_bounds__(32 * kFusedWarps, kFusedCtasPerSm)
stft_pit_fused_kernel(const float* __restrict__ mixture, const float* __restrict__ yabs,
                      const float* __restrict__ sources, const float* __restrict__ mask,
                      const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames,
                      int shift, int64_t pad_left, const float* __restrict__ win,
                      const float2* __restrict__ twtab, int64_t gpe, int slots,
                      double* __restrict__ partial, int* __restrict__ counters, float* __restrict__ loss,
                      int32_t* __restrict__ perm, double* __restrict__ sse) {
  constexpr int NV = K * K;
  constexpr int F = fft::kBins;
  extern __shared__ __align__(16) float stage[];   // [2][rows][span] signal rows (rows = K, +1 when |Y| is
                                                   // recomputed), then the warps' mask / |Y| areas
  __shared__ float2 tiles[kFusedWarps][fft::kTile];
  constexpr int kYOffset = ((K * F + 3) / 4) * 4;            // |Y| row after the K mask rows
  constexpr int kWarpArea = ((kYOffset + F + 3) / 4) * 4;    // floats per warp
  __shared__ double sm[NV * kFusedWarps + NV];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int span = (kFusedWarps - 1) * shift + fft::kSize;
  const int nrows = yabs ? K : K + 1;              // staged signal rows per group
  const int buf_floats = nrows * span;
  float* wmask = stage + 2 * buf_floats;            // [kFusedWarps][kWarpArea] (span is a multiple of 4)
  float2* tile = tiles[warp];
  fft::LaneConsts<false> k;
  k.init(twtab, lane);
  float2 wa[8], wb[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    wa[r] = reinterpret_cast<const float2*>(win)[fft::natural_a(lane, r)];
    wb[r] = reinterpret_cast<const float2*>(win)[fft::natural_b(lane, r)];
  }
  const int k0 = fft::bin_a(lane, 0), k4 = fft::bin_a(lane, 4) - 256;   // bin of slot p = (p<4 ? k0 : k4) + 64 p
  const bool dup = lane == 0;   // lane 0: slot 15 duplicates bin 256, slots 16/17 = DC/Nyquist are live

  const int64_t total = batch * gpe;
  const int64_t g_begin = range_start(blockIdx.x, total, gridDim.x);
  const int64_t g_end = range_start(blockIdx.x + 1, total, gridDim.x);

  // stage the signal rows of group g into buffer `which`
  auto stage_rows = [&](int64_t g, int which) {
    const int64_t b = g / gpe, m0 = (g - b * gpe) * kFusedWarps;
    const int64_t Tb = meta ? meta[2 * b] : samples;
    const int64_t s0 = m0 * shift - pad_left;
    float* buf = stage + which * buf_floats;
    for (int j = 0; j < nrows; ++j) {
      const float* xr = (j < K) ? sources + (b * K + j) * samples : mixture + b * samples;
      if (VEC16) {
        fft::stage_group(buf + j * span, xr, s0, span, Tb);
      } else {
        for (int c = threadIdx.x; c < span; c += blockDim.x) {   // unaligned rows: plain loads
          const int64_t i = s0 + c;
          buf[j * span + c] = (i >= 0 && i < Tb) ? __ldg(xr + i) : 0.f;
        }
      }
    }
  };

  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  // flush the CTA's partial sums of example b (all threads call it)
  auto flush = [&](int64_t b) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float s = warp_sum(acc[i]);
      if (lane == 0) sm[i * kFusedWarps + warp] = (double)s;
      acc[i] = 0.f;
    }
    __syncthreads();
    const int64_t first = owner_of(b * gpe, total, gridDim.x);
    const int64_t last = owner_of((b + 1) * gpe - 1, total, gridDim.x);
    const int slot = (int)(blockIdx.x - first), nparts = (int)(last - first + 1);
    double* mine = partial + (b * slots + slot) * NV;
    if (threadIdx.x < NV) {
      double s = 0.0;
      for (int w = 0; w < kFusedWarps; ++w) s += sm[threadIdx.x * kFusedWarps + w];
      mine[threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counters + b, 1) == nparts - 1;
    __syncthreads();
    if (s_last) {   // block-uniform
      __threadfence();
      double* totals = sm + NV * kFusedWarps;
      if (threadIdx.x < NV) {
        double s = 0.0;
        const volatile double* p = partial + b * slots * NV + threadIdx.x;
        for (int c = 0; c < nparts; ++c) s += p[(int64_t)c * NV];
        totals[threadIdx.x] = s;
        sse[b * NV + threadIdx.x] = s;
      }
      __syncthreads();
      double best;
      int bp[B2S_MAX_SOURCES];
      search_permutations(totals, K, best, bp);
      if (threadIdx.x == 0) {
        const int64_t Mb = meta ? meta[2 * b + 1] : frames;
        loss[b] = (float)(best / ((double)Mb * (double)K * (double)F));
        for (int kk = 0; kk < K; ++kk) perm[b * K + kk] = bp[kk];
        counters[b] = 0;
      }
    }
    __syncthreads();
  };

  if (g_begin < g_end) stage_rows(g_begin, 0);
  fft::cp_async_commit();
  int cur = 0;
  int64_t b_cur = g_begin < g_end ? g_begin / gpe : -1;
  for (int64_t g = g_begin; g < g_end; ++g, cur ^= 1) {
    const int64_t b = g / gpe, m0 = (g - b * gpe) * kFusedWarps;
    if (b != b_cur) {   // CTA-uniform
      flush(b_cur);
      b_cur = b;
    }
    fft::cp_async_wait_all();
    __syncthreads();   // group g staged and visible; the other buffer is free
    if (g + 1 < g_end) stage_rows(g + 1, cur ^ 1);
    fft::cp_async_commit();

    const int64_t Mb = meta ? meta[2 * b + 1] : frames;
    const int64_t m = m0 + warp;
    if (m < Mb) {
      // this warp's mask rows [K][F] and |Y| row [F] travel to its private shared-memory area while the
      // transforms run (mask rows of a frame are contiguous and 8-byte aligned; |Y| rows only 4-byte)
      float* wm = wmask + warp * kWarpArea;          // [K * F] masks, then [F] |Y| at kYOffset
      {
        const float* mrow = mask + ((b * frames + m) * K) * F;
        if (MASK8) {
          for (int c = lane; c < (K * F) / 2; c += 32) fft::cp_async_8(wm + 2 * c, mrow + 2 * c);
          if ((K * F) & 1) { if (lane == 0) fft::cp_async_4(wm + K * F - 1, mrow + K * F - 1); }
        } else {
          for (int c = lane; c < K * F; c += 32) fft::cp_async_4(wm + c, mrow + c);
        }
        if (yabs) {
          const float* row = yabs + (b * frames + m) * F;
          for (int c = lane; c < F; c += 32) fft::cp_async_4(wm + kYOffset + c, row + c);
        }
        fft::cp_async_commit();
      }
      const float* gbuf = stage + cur * buf_floats + warp * shift;
      float xs[K > 1 ? K - 1 : 1][18];   // magnitudes of the sources already transformed
      float x[18];
#pragma unroll 1
      for (int j = yabs ? 0 : -1; j < K; ++j) {
        const float2* src = reinterpret_cast<const float2*>(gbuf + (j < 0 ? K : j) * span);
        float2 a[8], bb[8], ya[8], yb[8];
        float ydc, ynyq;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 va = src[fft::natural_a(lane, r)], vb = src[fft::natural_b(lane, r)];
          a[r] = fft::pmul(va, wa[r]);
          bb[r] = fft::pmul(vb, wb[r]);
        }
        fft::rfft1024(a, bb, tile, k, ya, yb, ydc, ynyq);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          x[p] = fft::sqrt_approx(fmaf(ya[p].x, ya[p].x, ya[p].y * ya[p].y));
          x[8 + p] = fft::sqrt_approx(fmaf(yb[p].x, yb[p].x, yb[p].y * yb[p].y));
        }
        x[16] = fabsf(ydc); x[17] = fabsf(ynyq);
        if (j < 0) {
          // recomputed |Y|: park it in the warp's |Y| row so that the epilogue below is the same
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const int kk = (p < 4 ? k0 : k4) + 64 * p;
            wm[kYOffset + kk] = x[p];
            if (p < 7 || !dup) wm[kYOffset + fft::kHalf - kk] = x[8 + p];
          }
          if (dup) { wm[kYOffset] = x[16]; wm[kYOffset + fft::kHalf] = x[17]; }
        } else {
#pragma unroll
          for (int jj = 0; jj + 1 < K; ++jj)
            if (j == jj) {
#pragma unroll
              for (int q = 0; q < 18; ++q) xs[jj][q] = x[q];
            }
        }
      }
      // ---- SSE of this frame: e_i = mask_i * |Y| against every source magnitude
      fft::cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 18; ++q) {
        // bin of slot q; lanes >= 1 have no DC/Nyquist slots, lane 0 no second copy of bin 256
        int kk;
        bool live = true;
        if (q < 8) kk = (q < 4 ? k0 : k4) + 64 * q;
        else if (q < 16) { kk = fft::kHalf - ((q - 8 < 4 ? k0 : k4) + 64 * (q - 8)); live = q < 15 || !dup; }
        else { kk = q == 16 ? 0 : fft::kHalf; live = dup; }
        if (live) {
          const float o = wm[kYOffset + kk];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            const float e = wm[i * F + kk] * o;
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
              const float d = e - (jj + 1 < K ? xs[jj < K - 1 ? jj : 0][q] : x[q]);
              acc[i * K + jj] = fmaf(d, d, acc[i * K + jj]);
            }
          }
        }
      }
      __syncwarp();   // the area is rewritten by the next frame's copies
    }
  }
  fft::cp_async_wait_all();
  if (b_cur >= 0) flush(b_cur);
}

template <int K>
int launch_fused(const b2s_stft_plan* plan, const float* mixture, const float* yabs, const float* sources,
                 const float* mask, const int64_t* meta, int64_t batch, int64_t samples, int64_t frames,
                 int64_t pad_left, float* loss, int32_t* perm, double* sse, void* workspace,
                 cudaStream_t stream) {
  constexpr int sources_k_ = K;
  const FusedGrid g = fused_grid(batch, frames);
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = al16(sources) && (mixture == nullptr || al16(mixture)) && samples % 4 == 0 &&
                   plan->shift % 4 == 0 && pad_left % 4 == 0;
  const int span = (kFusedWarps - 1) * plan->shift + fft::kSize;
  const int nrows = yabs ? K : K + 1;
  const int warp_area = (((sources_k_ * fft::kBins + 3) / 4) * 4 + fft::kBins + 3) / 4 * 4;
  const size_t smem = sizeof(float) * (2 * nrows * span + kFusedWarps * warp_area);
  // mask rows [K][513] of a frame start at multiples of 8 * 513 * K / 2 bytes: 8-byte aligned with the base
  const bool mask8 = (reinterpret_cast<uintptr_t>(mask) & 7) == 0 && (K % 2 == 0);
  auto kernel = vec ? (mask8 ? stft_pit_fused_kernel<K, true, true> : stft_pit_fused_kernel<K, true, false>)
                    : (mask8 ? stft_pit_fused_kernel<K, false, true> : stft_pit_fused_kernel<K, false, false>);
  B2S_REQUIRE(smem <= 160 * 1024, "shift %d needs %zu bytes of staging: too large", plan->shift, smem);
  static bool configured[4][64] = {};   // per (variant, device)
  const int variant = (vec ? 2 : 0) + (mask8 ? 1 : 0);
  if (!configured[variant][plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured[variant][plan->device & 63] = true;
  }
  kernel<<<g.grid, 32 * kFusedWarps, smem, stream>>>(mixture, yabs, sources, mask, meta, batch, samples,
      frames, plan->shift, pad_left, plan->awin, plan->tw, g.gpe, g.slots, partial, counters, loss, perm, sse);
  B2S_LAUNCH_CHECK("stft_pit_fused_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int64_t b2s_stft_pit_workspace_bytes(int64_t batch, int64_t frames, int sources) {
  if (batch <= 0 || sources <= 0) return kTicketBytes + 16;
  const FusedGrid g = fused_grid(batch, frames);
  return kTicketBytes + (int64_t)sizeof(double) * batch * g.slots * sources * sources + 16;
}

int b2s_stft_pit_forward(const b2s_stft_plan* plan, const float* mixture, const float* observation_abs,
                         const float* sources, const float* mask, const int64_t* meta, int64_t batch,
                         int64_t samples, int sources_k, int64_t frames, int64_t pad_left, float* loss,
                         int32_t* perm, double* sse, void* workspace, b2s_stream stream) {
  B2S_REQUIRE(plan != nullptr, "stft plan is NULL");
  B2S_REQUIRE(plan->fast && plan->wlen == fft::kSize && plan->shift <= fft::kSize,
              "the fused STFT->PIT kernel exists for size 1024 / window_length 1024 / shift <= 1024 plans "
              "only (got size %d, window_length %d, shift %d)", plan->size, plan->wlen, plan->shift);
  B2S_REQUIRE(sources_k >= 1 && sources_k <= 4, "fused STFT->PIT supports 1..4 sources (got %d)", sources_k);
  B2S_REQUIRE(batch >= 0 && batch <= kMaxTickets && samples >= 0 && frames >= 0 && pad_left >= 0, "bad extents");
  B2S_REQUIRE(mixture || observation_abs, "need the mixture or its magnitude spectrogram");
  if (batch == 0) return B2S_OK;
  B2S_REQUIRE(sources && mask && loss && perm && sse && workspace, "NULL device pointer");
  B2S_CUDA(cudaSetDevice(plan->device));
  cudaStream_t st = (cudaStream_t)stream;
  switch (sources_k) {
    case 1: return launch_fused<1>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
    case 2: return launch_fused<2>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
    case 3: return launch_fused<3>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
    default: return launch_fused<4>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
  }
}

}  // extern "C"
