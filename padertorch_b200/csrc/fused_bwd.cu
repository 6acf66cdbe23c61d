// Backward of the fused north-star kernel (fused.cu): the gradient of the PIT-MSE loss w.r.t. the mask,
//     grad_mask[t, i, f] = g_b * 2 / (M_b K F) * (mask[t, i, f] |Y[t, f]| - |X_j(i)[t, f]|) * |Y[t, f]|,
// j(i) = the target the forward pass matched estimate i with (perm[j] == i), with the target spectra
// |X_k| = |STFT(s_k)| RECOMPUTED in registers from the waveforms -- like the forward pass they are never read from
// (or written to) HBM.  This is what autograd derives for pit_loss(mask * Y_abs[:, None, :], X_abs, axis=-2)
// (padertorch/contrib/examples/source_separation/pit/model.py:117-128 through
//  padertorch/ops/losses/source_separation.py:112-119 and torch.nn.functional.mse_loss).
// Algorithmic HBM bytes per utterance: 4T(1+K) (waveforms) + 4MFK (mask) + 4MFK (gradient)   [SURVEY.md 8d].
//
// Same warp pipelines as the forward kernel: every warp owns a contiguous range of frame positions, its elected
// lane starts TMA bulk copies of the K source frames and of the position's mask rows [K][513] / |Y| row; the
// position's gradient block [K][513] -- contiguous in global memory -- is assembled in shared memory at the
// destination's phase within 16 bytes and leaves as ONE asynchronous TMA bulk store.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "fft1024.cuh"
#include "rfft_packed.cuh"
#include "tma.cuh"
#include "stft_plan.cuh"

using namespace b2s;
using namespace b2s::tma;

namespace {

struct BwdShape { int warps, ctas; };
// shared memory per warp: NT frames + 2 exchange tiles + mask / |Y| rows + the gradient block
__host__ __device__ constexpr BwdShape bwd_shape(int K, bool recompute) {
  if (K <= 2) return recompute ? BwdShape{7, 1} : BwdShape{4, 2};
  if (K == 3) return recompute ? BwdShape{5, 1} : BwdShape{6, 1};
  return recompute ? BwdShape{4, 1} : BwdShape{5, 1};
}
__host__ __device__ constexpr int mask_floats(int K) { return ((K * 513 + 8 + 3) / 4) * 4; }
__host__ __device__ constexpr int yrow_floats() { return ((513 + 8 + 3) / 4) * 4; }
__host__ __device__ constexpr int bwd_warp_floats(int K, bool recompute) {
  return ((recompute ? K + 1 : K) * rf::kSize) + 4 * rf::kTile1 + mask_floats(K) + yrow_floats() + mask_floats(K);
}

__device__ __forceinline__ float2 mag2(float2 ya, float2 yb) {
  return make_float2(fft::sqrt_approx(fmaf(ya.x, ya.x, ya.y * ya.y)),
                     fft::sqrt_approx(fmaf(yb.x, yb.x, yb.y * yb.y)));
}

template <int K, bool RECOMPUTE_Y, int WARPS, int CTAS>
__global__ void __launch_bounds__(32 * WARPS, CTAS)
stft_pit_fused_backward_kernel(const float* __restrict__ mixture, const float* __restrict__ yabs,
                               const float* __restrict__ sources, const float* __restrict__ mask,
                               const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames,
                               int shift, int64_t pad_left, const float4* __restrict__ lane_table,
                               const int32_t* __restrict__ perm, const float* __restrict__ grad_loss,
                               float* __restrict__ grad_mask) {
  constexpr int F = rf::kBins;
  constexpr int NT = RECOMPUTE_Y ? K + 1 : K;
  constexpr int kWarpFloats = bwd_warp_floats(K, RECOMPUTE_Y);
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t bars[WARPS][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sig = smem + warp * kWarpFloats;
  float2* tile = reinterpret_cast<float2*>(sig + NT * rf::kSize);
  float* rows_area = sig + NT * rf::kSize + 4 * rf::kTile1;
  float* out_area = rows_area + mask_floats(K) + yrow_floats();
  uint64_t* bar_sig = &bars[warp][0];
  uint64_t* bar_rows = &bars[warp][1];
  if (lane == 0) {
    mbar_init(bar_sig, 1);
    mbar_init(bar_rows, 1);
    fence_mbar_init();
  }
  __syncwarp();
  rf::LaneConsts k;
  k.load(lane_table, lane);
  const int k0 = rf::bin_a(lane, 0), k4 = rf::bin_a(lane, 4) - 256;
  const bool first = lane == 0;

  const int64_t total = batch * frames;
  const int64_t nwarps = min((int64_t)gridDim.x * WARPS, total);
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
  if (gw >= nwarps) return;
  const int64_t p_begin = gw * total / nwarps, p_end = (gw + 1) * total / nwarps;

  unsigned sig_phase = 0, rows_phase = 0;
  bool sig_by_tma = false;
  int off_m = 0, off_y = 0;

  // per-example context (warp uniform), recomputed only when the example changes
  int64_t ctx_b = -1;
  int ctx_T = 0, ctx_M = 0;
  bool ctx_a16 = false;
  const float* ctx_row[NT];
  const float* ctx_mask = nullptr;
  const float* ctx_y = nullptr;
  float* ctx_grad = nullptr;
  float ctx_scale = 0.f;
  int ctx_match[K];   // target matched with estimate i
  auto set_ctx = [&](int64_t b) {
    if (b == ctx_b) return;
    ctx_b = b;
    ctx_T = (int)(meta ? meta[2 * b] : samples);
    ctx_M = (int)(meta ? meta[2 * b + 1] : frames);
    ctx_a16 = true;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int r = RECOMPUTE_Y ? (t == 0 ? K : t - 1) : t;
      ctx_row[t] = r < K ? sources + (b * K + r) * samples : mixture + b * samples;
      ctx_a16 = ctx_a16 && (reinterpret_cast<uintptr_t>(ctx_row[t]) & 15) == 0;
    }
    ctx_mask = mask + b * frames * (K * F);
    ctx_grad = grad_mask + b * frames * (K * F);
    ctx_y = RECOMPUTE_Y ? nullptr : yabs + b * frames * F;
    ctx_scale = grad_loss[b] * (2.f / ((float)ctx_M * (float)(K * F)));
#pragma unroll
    for (int i = 0; i < K; ++i) ctx_match[i] = i;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int i = perm[b * K + j];
#pragma unroll
      for (int ii = 0; ii < K; ++ii) if (ii == i) ctx_match[ii] = j;
    }
  };
  const int pad = (int)pad_left;
  auto start_signals = [&](int64_t q, int64_t b, int m) {
    if (q >= p_end) return;
    set_ctx(b);
    if (m >= ctx_M) { sig_by_tma = false; return; }
    const int s0 = m * shift - pad;
    const bool a16 = ctx_a16 && (s0 & 3) == 0;
    const bool bulk = a16 && s0 >= 0 && s0 + rf::kSize <= ctx_T;
    sig_by_tma = bulk;
    if (bulk) {
      if (lane == 0) {
        mbar_expect_tx(bar_sig, NT * rf::kSize * 4u);
#pragma unroll
        for (int t = 0; t < NT; ++t) bulk_g2s(sig + t * rf::kSize, ctx_row[t] + s0, rf::kSize * 4u, bar_sig);
      }
    } else {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const float* xr = ctx_row[t];
        if (a16) {
          for (int c = lane; c < rf::kSize / 4; c += 32) {
            const int n = s0 + 4 * c;
            const int bytes = n < 0 ? 0 : max(0, min(4, ctx_T - n)) * 4;
            fft::cp_async_16(sig + t * rf::kSize + 4 * c, bytes ? xr + n : xr, bytes);
          }
        } else {
          for (int i = lane; i < rf::kSize; i += 32) {
            const int n = s0 + i;
            const bool ok = n >= 0 && n < ctx_T;
            fft::cp_async_4_zfill(sig + t * rf::kSize + i, ok ? xr + n : xr, ok ? 4 : 0);
          }
        }
      }
      fft::cp_async_commit();
    }
  };
  auto start_rows = [&](int64_t q, int64_t b, int m) {
    if (q >= p_end) return;
    set_ctx(b);
    if (m >= ctx_M) return;
    const uintptr_t am = reinterpret_cast<uintptr_t>(ctx_mask + m * (K * F));
    const uintptr_t ay = RECOMPUTE_Y ? 0 : reinterpret_cast<uintptr_t>(ctx_y + m * F);
    off_m = (int)(am & 15) >> 2;
    off_y = (int)(ay & 15) >> 2;
    if (lane == 0) {
      const unsigned bytes_m = (unsigned)(((am & 15) + K * F * 4 + 15) & ~15u);
      const unsigned bytes_y = RECOMPUTE_Y ? 0u : (unsigned)(((ay & 15) + F * 4 + 15) & ~15u);
      mbar_expect_tx(bar_rows, bytes_m + bytes_y);
      bulk_g2s(rows_area, reinterpret_cast<const void*>(am & ~(uintptr_t)15), bytes_m, bar_rows);
      if (!RECOMPUTE_Y)
        bulk_g2s(rows_area + mask_floats(K), reinterpret_cast<const void*>(ay & ~(uintptr_t)15), bytes_y, bar_rows);
    }
  };

  int64_t b = p_begin / frames;
  int m = (int)(p_begin - b * frames);
  const int frames_i = (int)frames;
  // nothing of the caller's tensors is requested before the preceding kernel has finished (PDL launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  start_signals(p_begin, b, m);
  start_rows(p_begin, b, m);
  for (int64_t q = p_begin; q < p_end; ++q) {
    int64_t bn = b;
    int mn = m + 1;
    if (mn == frames_i) { mn = 0; ++bn; }
    set_ctx(b);
    if (m >= ctx_M) {   // beyond this example's length (ragged batch): the gradient there stays zero
      start_signals(q + 1, bn, mn);
      start_rows(q + 1, bn, mn);
      b = bn; m = mn;
      continue;
    }
    // this position's context: the copies of the NEXT position (issued below) may switch the example
    const float scale = ctx_scale;
    float* const gdst = ctx_grad + (int64_t)m * (K * F);
    int match[K];
#pragma unroll
    for (int i = 0; i < K; ++i) match[i] = ctx_match[i];
    if (sig_by_tma) {
      mbar_wait(bar_sig, sig_phase);
      sig_phase ^= 1;
    } else {
      fft::cp_async_wait_all();
      __syncwarp();
    }
    auto magnitudes = [&](const float2 (&ya)[8], const float2 (&yb)[8], float ydc, float ynyq, float2 (&x)[9]) {
#pragma unroll
      for (int p = 0; p < 8; ++p) x[p] = mag2(ya[p], yb[p]);
      if (first) x[7].y = 0.f;
      x[8] = first ? make_float2(fabsf(ydc), fabsf(ynyq)) : make_float2(0.f, 0.f);
    };
    float2 x[NT][9];
#pragma unroll(K <= 2 ? 2 : 1)
    for (int t = 0; t + 1 < NT; t += 2) {
      float2 ya[2][8], yb[2][8];
      float ydc[2], ynyq[2];
      auto next_copy = [&]() { if (t + 2 >= NT) start_signals(q + 1, bn, mn); };
      rf::rfft_streams<2, false, false>(sig + t * rf::kSize, rf::kSize, tile, k, ya, yb, ydc, ynyq, 0, next_copy);
      magnitudes(ya[0], yb[0], ydc[0], ynyq[0], x[t]);
      magnitudes(ya[1], yb[1], ydc[1], ynyq[1], x[t + 1]);
    }
    if (NT & 1) {
      float2 ya[1][8], yb[1][8];
      float ydc[1], ynyq[1];
      auto next_copy = [&]() { start_signals(q + 1, bn, mn); };
      rf::rfft_streams<1, false, false>(sig + (NT - 1) * rf::kSize, 0, tile, k, ya, yb, ydc, ynyq, 0, next_copy);
      magnitudes(ya[0], yb[0], ydc[0], ynyq[0], x[NT - 1]);
    }
    constexpr int XS = RECOMPUTE_Y ? 1 : 0;

    mbar_wait(bar_rows, rows_phase);
    rows_phase ^= 1;
    const float* mrow = rows_area + off_m;
    const float* yrow = rows_area + mask_floats(K) + off_y;
    // the gradient block [K][F] is assembled at the destination's phase within 16 bytes
    const int phase = (int)((reinterpret_cast<uintptr_t>(gdst) & 15) >> 2);
    float* obuf = out_area + phase;
    if (lane == 0) bulk_wait_read<0>();   // the previous position's store has read the staging block
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 9; ++p) {
      const int ka = p < 8 ? (p < 4 ? k0 : k4) + 64 * p : 0;
      const int kb = rf::kHalf - ka;
      const bool live_a = p < 8 || first, live_b = p < 7 || (p == 7 ? !first : first);
      float2 ov;
      if (!RECOMPUTE_Y) ov = make_float2(live_a ? yrow[ka] : 0.f, live_b ? yrow[kb] : 0.f);
      else ov = x[0][p];
      const float2 ovs = rf::mul2(ov, rf::bcast(scale));
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float2 mv = make_float2(live_a ? mrow[i * F + ka] : 0.f, live_b ? mrow[i * F + kb] : 0.f);
        float2 xt = x[XS][p];
#pragma unroll
        for (int j = 1; j < K; ++j) if (match[i] == j) xt = x[XS + j][p];
        // (mask |Y| - |X|) * |Y| * scale
        const float2 d = rf::mul2(rf::sub2(rf::mul2(mv, ov), xt), ovs);
        if (live_a) obuf[i * F + ka] = d.x;
        if (live_b) obuf[i * F + kb] = d.y;
      }
    }
    fence_proxy_async();   // the block was written through the generic proxy
    __syncwarp();          // ... by every lane; every lane has also read its rows: the area may be overwritten
    constexpr int n = K * F;
    const int head = (4 - phase) & 3, mid = (n - head) & ~3, tail = n - head - mid;
    if (lane == 0) {
      bulk_s2g(gdst + head, obuf + head, (unsigned)mid * 4u);
      bulk_commit();
    }
    if (lane < head) gdst[lane] = obuf[lane];
    if (lane < tail) gdst[head + mid + lane] = obuf[head + mid + lane];
    start_rows(q + 1, bn, mn);
    b = bn; m = mn;
  }
  if (lane == 0) bulk_wait<0>();   // shared memory must outlive the last store
}

template <int K, bool RECOMPUTE>
int launch_backward(const b2s_stft_plan* plan, const float* mixture, const float* yabs, const float* sources,
                    const float* mask, const int64_t* meta, int64_t batch, int64_t samples, int64_t frames,
                    int64_t pad_left, const int32_t* perm, const float* grad_loss, float* grad_mask,
                    cudaStream_t stream) {
  constexpr BwdShape shape = bwd_shape(K, RECOMPUTE);
  constexpr size_t smem = sizeof(float) * shape.warps * bwd_warp_floats(K, RECOMPUTE);
  static_assert(smem <= 227 * 1024, "pipeline shape exceeds the shared memory of an SM");
  const int64_t total = batch * frames;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, shape.warps), (int64_t)kNumSMs * shape.ctas));
  auto kernel = stft_pit_fused_backward_kernel<K, RECOMPUTE, shape.warps, shape.ctas>;
  static bool configured[64] = {};
  if (!configured[plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[plan->device & 63] = true;
  }
  static const bool use_pdl = [] { const char* e = getenv("B2S_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * shape.warps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  const int shift = plan->shift;
  const float4* table = plan->lane_fwd;
  B2S_CUDA(cudaLaunchKernelEx(&cfg, kernel, mixture, yabs, sources, mask, meta, batch, samples, frames, shift,
                              pad_left, table, perm, grad_loss, grad_mask));
  B2S_LAUNCH_CHECK("stft_pit_fused_backward_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int b2s_stft_pit_backward(const b2s_stft_plan* plan, const float* mixture, const float* observation_abs,
                          const float* sources, const float* mask, const int64_t* meta, int64_t batch,
                          int64_t samples, int sources_k, int64_t frames, int64_t pad_left, const int32_t* perm,
                          const float* grad_loss, float* grad_mask, b2s_stream stream) {
  B2S_REQUIRE(plan != nullptr, "stft plan is NULL");
  B2S_REQUIRE(plan->fast && plan->wlen == fft::kSize && plan->shift <= fft::kSize && plan->shift % 4 == 0,
              "the fused STFT->PIT kernels exist for size 1024 / window_length 1024 / shift %% 4 == 0 plans only "
              "(got size %d, window_length %d, shift %d)", plan->size, plan->wlen, plan->shift);
  B2S_REQUIRE(sources_k >= 1 && sources_k <= 4, "fused STFT->PIT supports 1..4 sources (got %d)", sources_k);
  B2S_REQUIRE(batch >= 0 && samples >= 0 && frames >= 0 && pad_left >= 0, "bad extents");
  B2S_REQUIRE(samples < ((int64_t)1 << 30) && frames < ((int64_t)1 << 20) && pad_left < ((int64_t)1 << 30),
              "signal too long for the fused STFT->PIT kernels (%lld samples)", (long long)samples);
  B2S_REQUIRE(mixture || observation_abs, "need the mixture or its magnitude spectrogram");
  if (batch * frames == 0) return B2S_OK;
  B2S_REQUIRE(sources && mask && perm && grad_loss && grad_mask, "NULL device pointer");
  B2S_ON_DEVICE(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
#define B2S_BWD(K)                                                                                             \
  return observation_abs                                                                                       \
      ? launch_backward<K, false>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, \
                                  pad_left, perm, grad_loss, grad_mask, st)                                    \
      : launch_backward<K, true>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames,  \
                                 pad_left, perm, grad_loss, grad_mask, st)
  switch (sources_k) {
    case 1: B2S_BWD(1);
    case 2: B2S_BWD(2);
    case 3: B2S_BWD(3);
    default: B2S_BWD(4);
  }
#undef B2S_BWD
}

}  // extern "C"
