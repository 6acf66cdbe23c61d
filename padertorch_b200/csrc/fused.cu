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
#include <stdlib.h>

#include "common.cuh"
#include "fft1024.cuh"
#include "rfft_packed.cuh"
#include "tma.cuh"
#include "stft_plan.cuh"
#include "perm.cuh"

using namespace b2s;

namespace {

constexpr int kFusedWarps = 4;

struct FusedGrid {
  int grid;            // persistent CTAs
  int64_t gpe;         // groups per example (dense enumeration over `frames`)
  int64_t total;       // batch * gpe
  int slots;           // partial-sum slots per example
};

int sources_ctas(int K) {
  static const int want_ctas = [] { const char* e = getenv("B2S_FUSED_CTAS"); return e ? atoi(e) : 2; }();
  return (K <= 2 && want_ctas == 3) ? 3 : 2;
}

FusedGrid fused_grid(int64_t batch, int64_t frames, int ctas_per_sm) {
  FusedGrid g;
  g.gpe = std::max<int64_t>(1, ceil_div(frames, kFusedWarps));
  g.total = batch * g.gpe;
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>(g.total, (int64_t)kNumSMs * ctas_per_sm));
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

using namespace b2s::tma;

// floats of a warp's mask / |Y| landing area: [mask rows K * F + 8 | |Y| row F + 8], both 16-byte aligned
__host__ __device__ constexpr int mask_area_floats(int K) { return ((K * 513 + 8 + 3) / 4) * 4; }
__host__ __device__ constexpr int row_area_floats(int K) { return mask_area_floats(K) + ((513 + 8 + 3) / 4) * 4; }

// magnitudes of one slot pair (A side, B side)
__device__ __forceinline__ float2 mag2(float2 ya, float2 yb) {
  return make_float2(fft::sqrt_approx(fmaf(ya.x, ya.x, ya.y * ya.y)),
                     fft::sqrt_approx(fmaf(yb.x, yb.x, yb.y * yb.y)));
}

// One frame position = (K [+1]) transforms (rfft_packed.cuh) whose magnitudes stay in registers as packed
// (A side, B side) pairs, then the K x K SSE of 9 packed bin pairs per lane (slot 8 = DC / Nyquist, live in
// lane 0 only); mask and |Y| values are read straight from global memory at their point of use (the rows
// were prefetched into L2 one group earlier), each exactly once.
template <int K, bool VEC16, bool RECOMPUTE_Y, int CTAS>
__global__ void __launch_bounds__(32 * kFusedWarps, CTAS)
stft_pit_fused_kernel(const float* __restrict__ mixture, const float* __restrict__ yabs,
                      const float* __restrict__ sources, const float* __restrict__ mask,
                      const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames,
                      int shift, int64_t pad_left, const float4* __restrict__ lane_table, int64_t gpe, int slots,
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
  float2* tile = reinterpret_cast<float2*>(smem + 2 * buf_floats) + warp * (2 * rf::kTile1);
  // per warp: the mask rows [K][F] (+ |Y| row) of its current frame, copied by TMA while the transforms run
  float* rows_area = smem + 2 * buf_floats + kFusedWarps * 4 * rf::kTile1 + warp * row_area_floats(K);
  __shared__ __align__(8) uint64_t bars[kFusedWarps];
  uint64_t* bar = &bars[warp];
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  unsigned bar_phase = 0;       // parity of the copy the warp waits for next
  int off_m = 0, off_y = 0;     // float offset of the row inside its landing area (source misalignment / 4)
  bool copy_pending = false;
  rf::LaneConsts k;
  k.load(lane_table, lane);
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
  // Start the copy of the mask rows [K][F] (contiguous) and the |Y| row of this warp's frame of group g.  A bulk
  // copy needs 16-byte aligned addresses and sizes while rows start at multiples of 4 bytes: the enclosing
  // aligned range is copied and the row found at the source's misalignment inside the landing area.  The at
  // most 12 bytes before / after a row belong to the neighbouring rows or, at the very ends of the tensor, to
  // the same (>= 256-byte granular) allocation -- an address that is not 16-byte aligned is never at its edge.
  auto start_copy = [&](int64_t g) {
    const int64_t b = g / gpe, m = (g - b * gpe) * kFusedWarps + warp;
    const int64_t Mb = meta ? meta[2 * b + 1] : frames;
    copy_pending = m < Mb;
    if (!copy_pending) return;
    const uintptr_t am = reinterpret_cast<uintptr_t>(mask + ((b * frames + m) * K) * F);
    const uintptr_t ay = RECOMPUTE_Y ? 0 : reinterpret_cast<uintptr_t>(yabs + (b * frames + m) * F);
    off_m = (int)(am & 15) >> 2;
    off_y = (int)(ay & 15) >> 2;
    if (lane == 0) {
      const unsigned bytes_m = (unsigned)(((am & 15) + K * F * 4 + 15) & ~15u);
      const unsigned bytes_y = RECOMPUTE_Y ? 0u : (unsigned)(((ay & 15) + F * 4 + 15) & ~15u);
      fence_proxy_async();   // the area's previous contents were read through the generic proxy
      mbar_expect_tx(bar, bytes_m + bytes_y);
      bulk_g2s(rows_area, reinterpret_cast<const void*>(am & ~(uintptr_t)15), bytes_m, bar);
      if (!RECOMPUTE_Y)
        bulk_g2s(rows_area + mask_area_floats(K), reinterpret_cast<const void*>(ay & ~(uintptr_t)15), bytes_y, bar);
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
    start_copy(g_begin);
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
    }
    fft::cp_async_commit();

    const int64_t Mb = meta ? meta[2 * b + 1] : frames;
    const int64_t m = m0 + warp;
    if (m < Mb) {
      const float* gbuf = smem + cur * buf_floats + warp * shift;
      // transforms run two at a time (rfft_streams: the sources of a frame position are independent streams);
      // magnitudes stay in registers as (A side, B side) pairs; slot 8 = (|DC|, |Nyquist|), zero outside lane 0;
      // lane 0's slot 7 holds bin 256 on both sides: its B copy is zeroed (and so is the matching mask load)
      auto magnitudes = [&](const float2 (&ya)[8], const float2 (&yb)[8], float ydc, float ynyq, float2 (&x)[9]) {
#pragma unroll
        for (int p = 0; p < 8; ++p) x[p] = mag2(ya[p], yb[p]);
        if (first) x[7].y = 0.f;
        x[8] = first ? make_float2(fabsf(ydc), fabsf(ynyq)) : make_float2(0.f, 0.f);
      };
      // signal row r of the staged group: r < K sources, r == K the mixture; row order of the transforms:
      // (mixture,) source 0, source 1, ...
      constexpr int NT = RECOMPUTE_Y ? K + 1 : K;
      float2 x[NT][9];   // with RECOMPUTE_Y x[0] = |Y|, sources follow
      auto row_of = [&](int t) { return RECOMPUTE_Y ? (t == 0 ? K : t - 1) : t; };
#pragma unroll(K <= 2 ? 2 : 1)
      for (int t = 0; t + 1 < NT; t += 2) {
        float2 ya[2][8], yb[2][8];
        float ydc[2], ynyq[2];
        const int r0 = row_of(t), r1 = row_of(t + 1);
        rf::rfft_streams<2, false, false>(gbuf + r0 * span, (r1 - r0) * span, tile, k, ya, yb, ydc, ynyq);
        magnitudes(ya[0], yb[0], ydc[0], ynyq[0], x[t]);
        magnitudes(ya[1], yb[1], ydc[1], ynyq[1], x[t + 1]);
      }
      if (NT & 1) {
        float2 ya[1][8], yb[1][8];
        float ydc[1], ynyq[1];
        rf::rfft_streams<1, false, false>(gbuf + row_of(NT - 1) * span, 0, tile, k, ya, yb, ydc, ynyq);
        magnitudes(ya[0], yb[0], ydc[0], ynyq[0], x[NT - 1]);
      }
      constexpr int XS = RECOMPUTE_Y ? 1 : 0;   // x[XS + j] = |STFT(s_j)|

      // ---- SSE of this frame: e_i = mask_i * |Y| against every source magnitude
      mbar_wait(bar, bar_phase);   // the rows of this frame have landed
      bar_phase ^= 1;
      const float* mrow = rows_area + off_m;
      const float* yrow = rows_area + mask_area_floats(K) + off_y;
#pragma unroll
      for (int p = 0; p < 9; ++p) {
        const int ka = p < 8 ? (p < 4 ? k0 : k4) + 64 * p : 0;
        const int kb = rf::kHalf - ka;
        const bool live_a = p < 8 || first, live_b = p < 7 || (p == 7 ? !first : first);
        float2 ov;
        if (!RECOMPUTE_Y) ov = make_float2(live_a ? yrow[ka] : 0.f, live_b ? yrow[kb] : 0.f);
        else ov = x[0][p];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float2 mv = make_float2(live_a ? mrow[i * F + ka] : 0.f, live_b ? mrow[i * F + kb] : 0.f);
          const float2 e = rf::mul2(mv, ov);
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const float2 d = rf::sub2(e, x[XS + j][p]);
            acc[i * K + j] = rf::fma2(d, d, acc[i * K + j]);
          }
        }
      }
      __syncwarp();   // every lane has read its rows: the area may be overwritten
    }
    if (g + 1 < g_end) start_copy(g + 1);
  }
  fft::cp_async_wait_all();
  if (b_cur >= 0) flush(b_cur);
}

template <int K>
int launch_fused(const b2s_stft_plan* plan, const float* mixture, const float* yabs, const float* sources,
                 const float* mask, const int64_t* meta, int64_t batch, int64_t samples, int64_t frames,
                 int64_t pad_left, float* loss, int32_t* perm, double* sse, void* workspace,
                 cudaStream_t stream) {
  const FusedGrid g = fused_grid(batch, frames, sources_ctas(K));
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
  const size_t smem = sizeof(float) * 2 * nrows * span + sizeof(float2) * 2 * rf::kTile1 * kFusedWarps +
                      sizeof(float) * row_area_floats(K) * kFusedWarps;
  // resident CTAs per SM: 2 (<= 255 registers: the two interleaved transforms keep all their values in
  // registers); B2S_FUSED_CTAS=3 selects the 168-register build for experiments (shared memory must allow it)
  const bool three = sources_ctas(K) == 3;
  constexpr int C3 = K <= 2 ? 3 : 2;   // K >= 3 keeps more magnitudes live: always 2
  auto kernel = three
      ? (yabs ? (vec ? stft_pit_fused_kernel<K, true, false, C3> : stft_pit_fused_kernel<K, false, false, C3>)
              : (vec ? stft_pit_fused_kernel<K, true, true, C3> : stft_pit_fused_kernel<K, false, true, C3>))
      : (yabs ? (vec ? stft_pit_fused_kernel<K, true, false, 2> : stft_pit_fused_kernel<K, false, false, 2>)
              : (vec ? stft_pit_fused_kernel<K, true, true, 2> : stft_pit_fused_kernel<K, false, true, 2>));
  B2S_REQUIRE(smem <= 200 * 1024, "shift %d needs %zu bytes of staging: too large", plan->shift, smem);
  static bool configured[8][64] = {};   // per (variant, device)
  const int variant = (vec ? 1 : 0) + (yabs ? 0 : 2) + (three ? 4 : 0);
  if (!configured[variant][plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[variant][plan->device & 63] = true;
  }
  kernel<<<g.grid, 32 * kFusedWarps, smem, stream>>>(mixture, yabs, sources, mask, meta, batch, samples,
      frames, plan->shift, pad_left, plan->lane_fwd, g.gpe, g.slots, partial, counters, loss, perm, sse);
  B2S_LAUNCH_CHECK("stft_pit_fused_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int64_t b2s_stft_pit_workspace_bytes(int64_t batch, int64_t frames, int sources) {
  if (batch <= 0 || sources <= 0) return kTicketBytes + 16;
  const FusedGrid g = fused_grid(batch, frames, sources_ctas(sources));
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
