// Target / feature preparation of the PIT example straight from the waveforms:
//     Y_abs = |STFT(y)|,  X_abs = |STFT(s_k)|,  cos_phase_difference = cos(angle(Y) - angle(X_k))
// (pre_batch_transform, padertorch/contrib/examples/source_separation/pit/data.py:49-77) with the complex spectra
// never leaving the registers.  Same TMA-fed warp pipelines as the fused STFT -> PIT kernel (fused.cu): the frame
// positions of the whole batch form one list, every warp owns a contiguous range; per position the elected lane
// starts TMA bulk copies of the 1024-sample frames of the mixture and of the K sources, the K + 1 transforms run
// two at a time (rfft_packed.cuh), the unit phasor of Y stays in registers while the sources are transformed.
// The position's output rows -- |Y| [F], |X| [K][F] and the phase term [K][F], each contiguous in global memory
// ('t k f' layout) -- are assembled in shared memory at the destination's phase within 16 bytes and leave as three
// asynchronous TMA bulk stores (at most 3 floats at either end from lanes).
// Algorithmic HBM bytes per utterance: 4T(1 + K) read + 4MF(1 + 2K) written.
#include <algorithm>

#include "common.cuh"
#include "fft1024.cuh"
#include "rfft_packed.cuh"
#include "tma.cuh"
#include "stft_plan.cuh"

using namespace b2s;
using namespace b2s::tma;

namespace {

// warps per CTA: shared memory per warp (K + 1 frames, exchange tiles, 1 + 2K staged rows) grows with K; ONE CTA
// per SM with as many warps as fit in 227 KB (7 x 31.8 KB at K = 2; two CTAs of 3 warps would leave a seventh
// pipeline's worth of shared memory unused)
__host__ __device__ constexpr int tgt_warps(int K) { return K <= 1 ? 8 : (K == 2 ? 7 : (K == 3 ? 5 : 4)); }
constexpr int kTgtCtasPerSm = 1;
constexpr int F = rf::kBins;

__host__ __device__ constexpr int row_area(int rows) { return (rows * F + 3 + 3) / 4 * 4; }   // + misalignment
__host__ __device__ constexpr int tgt_warp_floats(int K) {
  return (K + 1) * rf::kSize + 4 * rf::kTile1 + row_area(1) + 2 * row_area(K);
}

// cos(angle(y) - angle(x)) from the unit phasor u of y (or 1 for y = 0) and x; numpy's angle(0) = 0
__device__ __forceinline__ float phase_term(float2 u, float2 x, float ax) {
  const float v = ax > 0.f ? __fdividef(fmaf(u.x, x.x, u.y * x.y), ax) : u.x;
  return fminf(1.f, fmaxf(-1.f, v));
}
__device__ __forceinline__ float2 unit_phasor(float2 y, float ay) {
  if (!(ay > 0.f)) return make_float2(1.f, 0.f);
  const float inv = __fdividef(1.f, ay);
  return make_float2(y.x * inv, y.y * inv);
}

template <int K>
__global__ void __launch_bounds__(32 * tgt_warps(K), kTgtCtasPerSm)
stft_targets_kernel(const float* __restrict__ mixture, const float* __restrict__ sources,
                    const int64_t* __restrict__ meta /* NULL or [B][2] = {samples_b, frames_b} */, int64_t batch,
                    int64_t samples, int64_t frames, int shift, int64_t pad_left,
                    const float4* __restrict__ lane_table, float* __restrict__ y_abs, float* __restrict__ x_abs,
                    float* __restrict__ cpd) {
  constexpr int NT = K + 1;   // transform 0: the mixture, 1..K: the sources
  constexpr int kTgtWarps = tgt_warps(K);
  constexpr int kWarpFloats = tgt_warp_floats(K);
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t bars[kTgtWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sig = smem + warp * kWarpFloats;
  float2* tile = reinterpret_cast<float2*>(sig + NT * rf::kSize);
  float* stage_y = sig + NT * rf::kSize + 4 * rf::kTile1;
  float* stage_x = stage_y + row_area(1);
  float* stage_c = stage_x + row_area(K);
  uint64_t* bar_sig = &bars[warp];
  if (lane == 0) {
    mbar_init(bar_sig, 1);
    fence_mbar_init();
  }
  __syncwarp();
  rf::LaneConsts k;
  k.load(lane_table, lane);
  // slot p holds bin (p < 4 ? k0 : k4) + 64 p on the A side and 512 minus that on the B side (rf::bin_a)
  const int k0 = rf::bin_a(lane, 0), k4 = rf::bin_a(lane, 4) - 256;
  const bool first = lane == 0;

  const int64_t total = batch * frames;
  const int64_t nwarps = min((int64_t)gridDim.x * kTgtWarps, total);
  const int64_t gw = (int64_t)blockIdx.x * kTgtWarps + warp;
  if (gw >= nwarps) return;
  const int64_t p_begin = gw * total / nwarps, p_end = (gw + 1) * total / nwarps;
  const int pad = (int)pad_left, frames_i = (int)frames;

  unsigned sig_phase = 0;
  bool sig_by_tma = false;
  int64_t ctx_b = -1;
  int T = (int)samples, ctx_M = frames_i;   // of the context's example (ragged batches: meta)
  bool ctx_a16 = false;
  const float* ctx_row[NT];
  auto set_ctx = [&](int64_t b) {
    if (b == ctx_b) return;
    ctx_b = b;
    if (meta) { T = (int)meta[2 * b]; ctx_M = (int)meta[2 * b + 1]; }
    ctx_a16 = true;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      ctx_row[t] = t == 0 ? mixture + b * samples : sources + (b * K + (t - 1)) * samples;
      ctx_a16 = ctx_a16 && (reinterpret_cast<uintptr_t>(ctx_row[t]) & 15) == 0;
    }
  };
  // start the copy of the NT frames of position (b, m): TMA, or zero-filling cp.async for frames that touch the
  // zero padding at the signal's ends or are not 16-byte aligned
  auto start_signals = [&](bool valid, int64_t b, int m) {
    if (!valid) return;
    set_ctx(b);
    if (m >= ctx_M) { sig_by_tma = false; return; }   // padding row of a shorter example: nothing to transform
    const int s0 = m * shift - pad;
    const bool a16 = ctx_a16 && (s0 & 3) == 0;
    const bool bulk = a16 && s0 >= 0 && s0 + rf::kSize <= T;
    sig_by_tma = bulk;
    if (bulk) {
      if (lane == 0) {
        // (no fence.proxy.async: the area was only READ through the generic proxy, every lane's loads were consumed
        // before the __syncwarp() that precedes this call, and the copy's writes arrive a memory latency later)
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
            const int bytes = n < 0 ? 0 : max(0, min(4, T - n)) * 4;
            fft::cp_async_16(sig + t * rf::kSize + 4 * c, bytes ? xr + n : xr, bytes);
          }
        } else {
          for (int i = lane; i < rf::kSize; i += 32) {
            const int n = s0 + i;
            const bool ok = n >= 0 && n < T;
            fft::cp_async_4_zfill(sig + t * rf::kSize + i, ok ? xr + n : xr, ok ? 4 : 0);
          }
        }
      }
      fft::cp_async_commit();
    }
  };

  int64_t b = p_begin / frames;
  int m = (int)(p_begin - b * frames);
  // launched with programmatic stream serialization: everything above is plan data or set-up and overlaps the
  // tail of the preceding kernel; no caller data is touched before this wait
  asm volatile("griddepcontrol.wait;" ::: "memory");
  start_signals(true, b, m);
  for (int64_t q = p_begin; q < p_end; ++q) {
    const bool has_next = q + 1 < p_end;
    int64_t bn = b;
    int mn = m + 1;
    if (mn == frames_i) { mn = 0; ++bn; }
    set_ctx(b);
    const bool live = m < ctx_M;
    if (sig_by_tma) {
      mbar_wait(bar_sig, sig_phase);
      sig_phase ^= 1;
    } else {
      fft::cp_async_wait_all();
      __syncwarp();
    }
    // destination rows of this position and their phases within 16 bytes
    const int64_t pos = b * frames + m;
    float* gy = y_abs + pos * F;
    float* gx = x_abs + pos * K * F;
    float* gc = cpd + pos * K * F;
    const int ph_y = (int)((reinterpret_cast<uintptr_t>(gy) & 15) >> 2);
    const int ph_x = (int)((reinterpret_cast<uintptr_t>(gx) & 15) >> 2);
    const int ph_c = (int)((reinterpret_cast<uintptr_t>(gc) & 15) >> 2);
    float* sy = stage_y + ph_y;
    float* sx = stage_x + ph_x;
    float* sc = stage_c + ph_c;

    float2 ua[8], ub[8];        // unit phasors of Y (A side, B side as delivered: conjugated, which the real part
    float sdc = 1.f, snyq = 1.f;   // of u conj(x) does not see); signs of the real DC / Nyquist bins
    bool staged_free = false;
    auto free_stage = [&]() {   // the previous position's stores have read the staging rows
      if (staged_free) return;
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      staged_free = true;
    };
    // rows of one transform: mixture (tr == 0) or source tr - 1
    auto emit = [&](int tr, const float2 (&ya)[8], const float2 (&yb)[8], float ydc, float ynyq) {
      free_stage();
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int ka = (p < 4 ? k0 : k4) + 64 * p, kb = rf::kHalf - ka;
        const float ma = fft::sqrt_approx(fmaf(ya[p].x, ya[p].x, ya[p].y * ya[p].y));
        const float mb = fft::sqrt_approx(fmaf(yb[p].x, yb[p].x, yb[p].y * yb[p].y));
        const bool live_b = p < 7 || !first;   // lane 0's slot 7 holds bin 256 on both sides
        if (tr == 0) {
          ua[p] = unit_phasor(ya[p], ma);
          ub[p] = unit_phasor(yb[p], mb);
          sy[ka] = ma;
          if (live_b) sy[kb] = mb;
        } else {
          float* rx = sx + (tr - 1) * F;
          float* rc = sc + (tr - 1) * F;
          rx[ka] = ma;
          rc[ka] = phase_term(ua[p], ya[p], ma);
          if (live_b) {
            rx[kb] = mb;
            rc[kb] = phase_term(ub[p], yb[p], mb);
          }
        }
      }
      if (first) {   // DC and Nyquist are real
        if (tr == 0) {
          sdc = ydc < 0.f ? -1.f : 1.f;
          snyq = ynyq < 0.f ? -1.f : 1.f;
          sy[0] = fabsf(ydc);
          sy[rf::kHalf] = fabsf(ynyq);
        } else {
          float* rx = sx + (tr - 1) * F;
          float* rc = sc + (tr - 1) * F;
          rx[0] = fabsf(ydc);
          rx[rf::kHalf] = fabsf(ynyq);
          rc[0] = ydc < 0.f ? -sdc : sdc;
          rc[rf::kHalf] = ynyq < 0.f ? -snyq : snyq;
        }
      }
    };
    if (!live) {   // a frame beyond this example's length: the padded tensors hold zeros there (as a collated batch)
      free_stage();
      for (int i = lane; i < F; i += 32) sy[i] = 0.f;
      for (int i = lane; i < K * F; i += 32) { sx[i] = 0.f; sc[i] = 0.f; }
      start_signals(has_next, bn, mn);
    }
#pragma unroll
    for (int t = 0; live && t + 1 < NT; t += 2) {
      float2 ya[2][8], yb[2][8];
      float ydc[2], ynyq[2];
      auto next_copy = [&]() { if (t + 2 >= NT) start_signals(has_next, bn, mn); };
      rf::rfft_streams<2, false, false>(sig + t * rf::kSize, rf::kSize, tile, k, ya, yb, ydc, ynyq, 0, next_copy);
      emit(t, ya[0], yb[0], ydc[0], ynyq[0]);
      emit(t + 1, ya[1], yb[1], ydc[1], ynyq[1]);
    }
    if (live && (NT & 1)) {
      float2 ya[1][8], yb[1][8];
      float ydc[1], ynyq[1];
      auto next_copy = [&]() { start_signals(has_next, bn, mn); };
      rf::rfft_streams<1, false, false>(sig + (NT - 1) * rf::kSize, 0, tile, k, ya, yb, ydc, ynyq, 0, next_copy);
      emit(NT - 1, ya[0], yb[0], ydc[0], ynyq[0]);
    }
    // ---- the staged rows leave: aligned middle by TMA, at most 3 floats at either end from lanes
    fence_proxy_async();   // the rows were written through the generic proxy
    __syncwarp();
    auto ship = [&](float* g, const float* st, int ph, int n) {
      const int head = (4 - ph) & 3, mid = (n - head) & ~3, tail = n - head - mid;
      if (lane == 0) bulk_s2g(g + head, st + head, (unsigned)mid * 4u);
      if (lane < head) g[lane] = st[lane];
      if (lane < tail) g[head + mid + lane] = st[head + mid + lane];
    };
    ship(gy, sy, ph_y, F);
    ship(gx, sx, ph_x, K * F);
    ship(gc, sc, ph_c, K * F);
    if (lane == 0) bulk_commit();
    b = bn; m = mn;
  }
  if (lane == 0) bulk_wait<0>();   // shared memory must outlive the last store
}

template <int K>
int launch_targets(const b2s_stft_plan* plan, const float* mixture, const float* sources, const int64_t* meta, int64_t batch,
                   int64_t samples, int64_t frames, int64_t pad_left, float* y_abs, float* x_abs, float* cpd,
                   cudaStream_t stream) {
  constexpr int kTgtWarps = tgt_warps(K);
  const int64_t total = batch * frames;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, kTgtWarps),
                                                               (int64_t)kNumSMs * kTgtCtasPerSm));
  const size_t smem = sizeof(float) * kTgtWarps * tgt_warp_floats(K);
  B2S_REQUIRE(smem <= 227 * 1024, "internal: %zu bytes of shared memory", smem);
  static bool configured[64] = {};
  if (!configured[plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(stft_targets_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[plan->device & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * kTgtWarps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int shift = plan->shift;
  const float4* table = plan->lane_fwd;
  B2S_CUDA(cudaLaunchKernelEx(&cfg, stft_targets_kernel<K>, mixture, sources, meta, batch, samples, frames, shift,
                              pad_left, table, y_abs, x_abs, cpd));
  B2S_LAUNCH_CHECK("stft_targets_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int b2s_stft_pit_targets(const b2s_stft_plan* plan, const float* mixture, const float* sources,
                         const int64_t* meta, int64_t batch, int64_t samples, int sources_k, int64_t frames, int64_t pad_left, float* y_abs,
                         float* x_abs, float* cos_phase_difference, b2s_stream stream) {
  B2S_REQUIRE(plan != nullptr, "stft plan is NULL");
  B2S_REQUIRE(plan->fast && plan->wlen == fft::kSize && plan->shift <= fft::kSize && plan->shift % 4 == 0,
              "the fused target preparation exists for size 1024 / window_length 1024 / shift %% 4 == 0 plans only "
              "(got size %d, window_length %d, shift %d)", plan->size, plan->wlen, plan->shift);
  B2S_REQUIRE(sources_k >= 1 && sources_k <= 4, "fused target preparation supports 1..4 sources (got %d)", sources_k);
  B2S_REQUIRE(batch >= 0 && samples >= 0 && frames >= 0 && pad_left >= 0, "bad extents");
  B2S_REQUIRE(samples < ((int64_t)1 << 30) && frames < ((int64_t)1 << 20) && pad_left < ((int64_t)1 << 30),
              "signal too long for the fused target preparation (%lld samples)", (long long)samples);
  if (batch * frames == 0) return B2S_OK;
  B2S_REQUIRE(mixture && sources && y_abs && x_abs && cos_phase_difference, "NULL device pointer");
  B2S_ON_DEVICE(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
  switch (sources_k) {
    case 1: return launch_targets<1>(plan, mixture, sources, meta, batch, samples, frames, pad_left, y_abs, x_abs, cos_phase_difference, st);
    case 2: return launch_targets<2>(plan, mixture, sources, meta, batch, samples, frames, pad_left, y_abs, x_abs, cos_phase_difference, st);
    case 3: return launch_targets<3>(plan, mixture, sources, meta, batch, samples, frames, pad_left, y_abs, x_abs, cos_phase_difference, st);
    default: return launch_targets<4>(plan, mixture, sources, meta, batch, samples, frames, pad_left, y_abs, x_abs, cos_phase_difference, st);
  }
}

}  // extern "C"
