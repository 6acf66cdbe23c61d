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
#include "rfft_packed.cuh"
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

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// magnitudes of one slot pair (A side, B side)
__device__ __forceinline__ float2 mag2(float2 ya, float2 yb) {
  return make_float2(fft::sqrt_approx(fmaf(ya.x, ya.x, ya.y * ya.y)),
                     fft::sqrt_approx(fmaf(yb.x, yb.x, yb.y * yb.y)));
}

// One frame position = (K [+1]) transforms (rfft_packed.cuh) whose magnitudes stay in registers as packed
// (A side, B side) pairs, then the K x K SSE of 9 packed bin pairs per lane (slot 8 = DC / Nyquist, live in
// lane 0 only); mask and |Y| values are read straight from global memory at their point of use (the rows
// were prefetched into L2 one group earlier), each exactly once.
template <int K, bool VEC16, bool RECOMPUTE_Y>
__global__ void __launch_bounds__(32 * kFusedWarps, K <= 2 ? kFusedCtasPerSm : 2)
stft_pit_fused_kernel(const float* __restrict__ mixture, const float* __restrict__ yabs,
                      const float* __restrict__ sources, const float* __restrict__ mask,
                      const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames,
                      int shift, int64_t pad_left, const float* __restrict__ win,
                      const float2* __restrict__ twtab, int64_t gpe, int slots,
                      double* __restrict__ partial, int* __restrict__ counters, float* __restrict__ loss,
                      int32_t* __restrict__ perm, double* __restrict__ sse) {
  constexpr int NV = K * K;
  constexpr int F = rf::kBins;
  extern __shared__ __align__(16) float smem[];   // [2][rows][span] signal rows (rows = K, +1 when |Y| is
                                                  // recomputed), then the warps' exchange tiles
  __shared__ double sm[NV * kFusedWarps + NV];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int span = (kFusedWarps - 1) * shift + rf::kSize;
  constexpr int nrows = RECOMPUTE_Y ? K + 1 : K;   // staged signal rows per group
  const int buf_floats = nrows * span;
  float2* tile = reinterpret_cast<float2*>(smem + 2 * buf_floats) + warp * rf::kTile;
  rf::LaneConsts k;
  k.init(twtab, win, lane);
  // slot p holds bin (p < 4 ? k0 : k4) + 64 p on the A side and 512 minus that on the B side
  const int k0 = rf::bin_a(lane, 0), k4 = rf::bin_a(lane, 4) - 256;
  const bool first = lane == 0;

  const int64_t total = batch * gpe;
  const int64_t g_begin = range_start(blockIdx.x, total, gridDim.x);
  const int64_t g_end = range_start(blockIdx.x + 1, total, gridDim.x);

  // stage the signal rows of group g into buffer `which`
  auto stage_rows = [&](int64_t g, int which) {
    const int64_t b = g / gpe, m0 = (g - b * gpe) * kFusedWarps;
    const int64_t Tb = meta ? meta[2 * b] : samples;
    const int64_t s0 = m0 * shift - pad_left;
    float* buf = smem + which * buf_floats;
    const bool interior = VEC16 && s0 >= 0 && s0 + span <= Tb;
    for (int j = 0; j < nrows; ++j) {
      const float* xr = (j < K) ? sources + (b * K + j) * samples : mixture + b * samples;
      if (interior) {
        for (int c = threadIdx.x; c < (span >> 2); c += blockDim.x)
          fft::cp_async_16(buf + j * span + 4 * c, xr + s0 + 4 * c, 16);
      } else if (VEC16) {
        fft::stage_group(buf + j * span, xr, s0, span, Tb);
      } else {
        for (int c = threadIdx.x; c < span; c += blockDim.x) {   // unaligned rows: plain loads
          const int64_t i = s0 + c;
          buf[j * span + c] = (i >= 0 && i < Tb) ? __ldg(xr + i) : 0.f;
        }
      }
    }
  };
  // pull the mask rows [K][F] (contiguous) and the |Y| row of this warp's frame of group g into L2
  auto prefetch_rows = [&](int64_t g) {
    const int64_t b = g / gpe, m = (g - b * gpe) * kFusedWarps + warp;
    const int64_t Mb = meta ? meta[2 * b + 1] : frames;
    if (m >= Mb) return;
    const char* mrow = reinterpret_cast<const char*>(mask + ((b * frames + m) * K) * F);
    for (int c = lane * 128; c < K * F * 4 + 127; c += 32 * 128) prefetch_l2(mrow + min(c, K * F * 4 - 4));
    if (!RECOMPUTE_Y) {
      const char* row = reinterpret_cast<const char*>(yabs + (b * frames + m) * F);
      if (lane * 128 < F * 4 + 127) prefetch_l2(row + min(lane * 128, F * 4 - 4));
    }
  };

  float2 acc[NV];   // (A-side sum, B-side sum)
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float2(0.f, 0.f);

  // flush the CTA's partial sums of example b (all threads call it)
  auto flush = [&](int64_t b) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float s = warp_sum(acc[i].x + acc[i].y);
      if (lane == 0) sm[i * kFusedWarps + warp] = (double)s;
      acc[i] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int64_t first_owner = owner_of(b * gpe, total, gridDim.x);
    const int64_t last_owner = owner_of((b + 1) * gpe - 1, total, gridDim.x);
    const int slot = (int)(blockIdx.x - first_owner), nparts = (int)(last_owner - first_owner + 1);
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

  if (g_begin < g_end) {
    stage_rows(g_begin, 0);
    prefetch_rows(g_begin);
  }
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
    if (g + 1 < g_end) {
      stage_rows(g + 1, cur ^ 1);
      prefetch_rows(g + 1);
    }
    fft::cp_async_commit();

    const int64_t Mb = meta ? meta[2 * b + 1] : frames;
    const int64_t m = m0 + warp;
    if (m < Mb) {
      const float* gbuf = smem + cur * buf_floats + warp * shift;
      // one transform: magnitudes as (A side, B side) pairs; slot 8 = (|DC|, |Nyquist|), zero outside lane 0;
      // lane 0's slot 7 holds bin 256 on both sides: its B copy is zeroed (and so is the matching mask load)
      auto transform = [&](const float* frame, float2 (&x)[9]) {
        float2 ya[8], yb[8];
        float ydc, ynyq;
        rf::pass1(frame, tile, k);
        __syncwarp();
        rf::pass2(tile, k);
        __syncwarp();
        rf::pass3(tile, k, ya, yb, ydc, ynyq);
#pragma unroll
        for (int p = 0; p < 8; ++p) x[p] = mag2(ya[p], yb[p]);
        if (first) x[7].y = 0.f;
        x[8] = first ? make_float2(fabsf(ydc), fabsf(ynyq)) : make_float2(0.f, 0.f);
      };
      float2 o[RECOMPUTE_Y ? 9 : 1];       // |Y| pairs when recomputed from the mixture
      if (RECOMPUTE_Y) transform(gbuf + K * span, reinterpret_cast<float2(&)[9]>(o));
      float2 x[K][9];
#pragma unroll(K <= 2 ? K : 1)
      for (int j = 0; j < K; ++j) transform(gbuf + j * span, x[j]);

      // ---- SSE of this frame: e_i = mask_i * |Y| against every source magnitude
      const float* mrow = mask + ((b * frames + m) * K) * F;
      const float* yrow = RECOMPUTE_Y ? nullptr : yabs + (b * frames + m) * F;
#pragma unroll
      for (int p = 0; p < 9; ++p) {
        const int ka = p < 8 ? (p < 4 ? k0 : k4) + 64 * p : 0;
        const int kb = rf::kHalf - ka;
        const bool live_a = p < 8 || first, live_b = p < 7 || (p == 7 ? !first : first);
        float2 ov;
        if (!RECOMPUTE_Y) ov = make_float2(live_a ? __ldg(yrow + ka) : 0.f, live_b ? __ldg(yrow + kb) : 0.f);
        else ov = o[RECOMPUTE_Y ? p : 0];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float2 mv = make_float2(live_a ? __ldg(mrow + i * F + ka) : 0.f,
                                        live_b ? __ldg(mrow + i * F + kb) : 0.f);
          const float2 e = rf::mul2(mv, ov);
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const float2 d = rf::sub2(e, x[j][p]);
            acc[i * K + j] = rf::fma2(d, d, acc[i * K + j]);
          }
        }
      }
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
  const FusedGrid g = fused_grid(batch, frames);
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = al16(sources) && (mixture == nullptr || al16(mixture)) && samples % 4 == 0 &&
                   plan->shift % 4 == 0 && pad_left % 4 == 0;
  // the transform reads its frame with 16-byte shared-memory loads: frames must start 16-byte aligned
  B2S_REQUIRE(plan->shift % 4 == 0, "the fused STFT->PIT kernel needs a shift that is a multiple of 4 (got %d)",
              plan->shift);
  const int span = (kFusedWarps - 1) * plan->shift + rf::kSize;
  const int nrows = yabs ? K : K + 1;
  const size_t smem = sizeof(float) * 2 * nrows * span + sizeof(float2) * rf::kTile * kFusedWarps;
  auto kernel = yabs ? (vec ? stft_pit_fused_kernel<K, true, false> : stft_pit_fused_kernel<K, false, false>)
                     : (vec ? stft_pit_fused_kernel<K, true, true> : stft_pit_fused_kernel<K, false, true>);
  B2S_REQUIRE(smem <= 200 * 1024, "shift %d needs %zu bytes of staging: too large", plan->shift, smem);
  static bool configured[4][64] = {};   // per (variant, device)
  const int variant = (vec ? 1 : 0) + (yabs ? 0 : 2);
  if (!configured[variant][plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
