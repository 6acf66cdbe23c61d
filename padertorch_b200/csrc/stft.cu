// STFT / iSTFT kernels and their C ABI (include/b200sep.h).
//
// Two transform engines behind one plan:
//   * size == 1024: warp-per-frame register FFT (fft1024.cuh); forward fuses zero padding, window,
//     FFT, real split and the magnitude / log1p epilogue; inverse fuses Hermitian extension, inverse
//     FFT, synthesis window and overlap-add in shared memory (halo frames recomputed per CTA so no
//     floating-point atomics are needed).
//   * every other size (any even size, any window_length <= size, any shift): table-driven DFT,
//     O(window_length * bins) per frame -- cheap for the short windows of TasNet's StftEncoder
//     (padertorch/contrib/examples/source_separation/tasnet/tas_coders.py:170) and exact for
//     non-power-of-two sizes.
// Reference arithmetic: padertorch/ops/_stft.py:103-263.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fft1024.cuh"
#include "rfft_packed.cuh"
#include "cfft_pair.cuh"
#include "stft_plan.cuh"
#include "tma.cuh"

using namespace b2s;


namespace {

// ------------------------------------------------------------------------------------------- epilogue
__device__ __forceinline__ void store_bin(float* __restrict__ out, int64_t frame, int k, float2 y,
                                          int layout, int bins) {
  if (layout == B2S_SPEC_INTERLEAVED) {
    reinterpret_cast<float2*>(out)[frame * bins + k] = y;
  } else if (layout == B2S_SPEC_CONCAT) {
    out[frame * 2 * bins + k] = y.x;
    out[frame * 2 * bins + bins + k] = y.y;
  } else {
    float m = sqrtf(fmaf(y.x, y.x, y.y * y.y));
    out[frame * bins + k] = layout == B2S_SPEC_ABS ? m : log1pf(m);
  }
}

__device__ __forceinline__ float2 load_bin(const float* __restrict__ in, int64_t frame, int k,
                                           int layout, int bins) {
  if (layout == B2S_SPEC_INTERLEAVED) return reinterpret_cast<const float2*>(in)[frame * bins + k];
  return make_float2(in[frame * 2 * bins + k], in[frame * 2 * bins + bins + k]);
}

// ------------------------------------------------------------------------------------------- fast forward
template <int LAYOUT>
__device__ __forceinline__ void store_bin_t(float* __restrict__ out, int k, float2 y) {
  // `out` already points at the frame's first bin
  if (LAYOUT == B2S_SPEC_INTERLEAVED) {
    reinterpret_cast<float2*>(out)[k] = y;
  } else if (LAYOUT == B2S_SPEC_CONCAT) {
    out[k] = y.x;
    out[fft::kBins + k] = y.y;
  } else {
    const float m = fft::sqrt_approx(fmaf(y.x, y.x, y.y * y.y));
    // log1p(m) = log(1 + m) through MUFU.LG2: absolute error <= ~3e-7 (rounding of 1 + m, 2^-22 of lg2.approx),
    // against a tolerance of 1e-4 x the utterance's largest value; libm's log1pf costs ~25 instructions more
    out[k] = LAYOUT == B2S_SPEC_ABS ? m : __logf(1.f + m);
  }
}

// Feature epilogue (layout B2S_SPEC_FEATURE): to_spectrogram + Logarithm + MelTransform of
// padertorch/contrib/mk/modules/features/timefreq.py:37-77, 171-183, 398-470 behind the transform.
struct FeatureArgs {
  float half_power;   // |Y|^power = (re^2 + im^2)^(power / 2)
  float scale;        // scale_spec: 1 / size, else 1
  int log_kind;       // B2S_LOG_*
  float eps;          // Logarithm: log(max(eps, x))
  int filters;        // 0: no filterbank
  const int* lo;      // [filters] first bin of filter m
  const int* len;     // [filters] number of bins of its support
  const int* woff;    // [filters] offset of its weights in `weights`
  const float* weights;
};
__device__ __forceinline__ float feature_power(float2 y, const FeatureArgs& a) {
  const float v = fmaf(y.x, y.x, y.y * y.y);
  float r;
  if (a.half_power == 0.5f) r = fft::sqrt_approx(v);
  else if (a.half_power == 1.f) r = v;
  else r = v > 0.f ? __powf(v, a.half_power) : 0.f;
  return r * a.scale;
}
__device__ __forceinline__ float feature_log(float x, const FeatureArgs& a) {
  if (a.log_kind == B2S_LOG_NONE) return x;
  x = fmaxf(x, a.eps);
  // lg2.approx: <= 2^-22 relative to |log2 x| plus the rounding of the factor
  const float l2 = __log2f(x);
  return a.log_kind == B2S_LOG_2 ? l2 : l2 * (a.log_kind == B2S_LOG_E ? 0.69314718055994530942f : 0.30102999566398119521f);
}

constexpr int kFwdWarps = 4;

// One warp per frame, persistent over frames.  `win` is the (zero-extended) window the samples are
// multiplied with: the analysis window for STFT, the synthesis window for the adjoint of iSTFT, in
// which case interior bins are doubled (`interior_scale` = 2).
template <int LAYOUT, bool VEC>
__global__ void __launch_bounds__(32 * kFwdWarps, 3)
stft1024_forward_kernel(const float* __restrict__ x, int64_t rows, int64_t samples, int64_t row_stride,
                        int64_t pad_left, int64_t frames, int shift, int wlen,
                        const float* __restrict__ win, const float2* __restrict__ twtab,
                        float interior_scale, float* __restrict__ out) {
  __shared__ float2 tiles[kFwdWarps][fft::kTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* tile = tiles[warp];
  fft::LaneConsts<false> k;
  k.init(twtab, lane);
  float2 wa[8], wb[8];   // window pairs of this lane's 16 packed samples
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    wa[r] = reinterpret_cast<const float2*>(win)[fft::natural_a(lane, r)];
    wb[r] = reinterpret_cast<const float2*>(win)[fft::natural_b(lane, r)];
  }
  constexpr int kOutPerFrame = LAYOUT <= B2S_SPEC_CONCAT ? 2 * fft::kBins : fft::kBins;
  const int64_t total = rows * frames;
  for (int64_t fr = (int64_t)blockIdx.x * kFwdWarps + warp; fr < total;
       fr += (int64_t)gridDim.x * kFwdWarps) {
    const int64_t row = fr / frames, m = fr - row * frames;
    float2 a[8], b[8], ya[8], yb[8];
    float ydc, ynyq;
    fft::load_frame<VEC>(x + row * row_stride, m * shift - pad_left, samples, wlen, lane, a, b);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a[r] = fft::pmul(a[r], wa[r]); b[r] = fft::pmul(b[r], wb[r]);
    }
    fft::rfft1024(a, b, tile, k, ya, yb, ydc, ynyq);
    float* o = out + fr * kOutPerFrame;
    const int k0 = fft::bin_a(lane, 0), k4 = fft::bin_a(lane, 4);   // slots 0..3 / 4..7 step by 64 bins
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int kk = (p < 4 ? k0 : k4 - 256) + 64 * p;
      float2 u = ya[p], v = yb[p];
      if (interior_scale != 1.f) {
        u.x *= interior_scale; u.y *= interior_scale; v.x *= interior_scale; v.y *= interior_scale;
      }
      store_bin_t<LAYOUT>(o, kk, u);
      if (p < 7 || lane != 0) store_bin_t<LAYOUT>(o, fft::kHalf - kk, v);
    }
    if (lane == 0) {
      store_bin_t<LAYOUT>(o, 0, make_float2(ydc, 0.f));
      store_bin_t<LAYOUT>(o, fft::kHalf, make_float2(ynyq, 0.f));
    }
  }
}

// ------------------------------------------------------------------------------------------- warp pipelines
// The default fast forward path.  Every warp is an independent pipeline: it owns units of two consecutive
// frames of one row, a TMA bulk copy (one instruction of one lane) brings the unit's shift + 1024 samples into
// the warp's own ring slot and signals an mbarrier, the copy of the next unit is started as soon as pass 1 has
// read the current one.  No block-wide barrier, no per-thread staging code: warps drift apart, so the copies,
// the shared-memory exchanges, the arithmetic and the spectrum stores of different warps overlap instead of
// marching through the same phase together (the block-synchronous staged kernel above measures T = T_copy +
// T_fft + T_store on B200, tools/ubench and B2S_ABLATE).  Units that touch the zero padding (fading, tail) or
// are not 16-byte aligned are filled by the warp itself.
constexpr int kPipeWarps = 4;      // warps per CTA
constexpr int kPipeCtasPerSm = 2;  // <= 255 registers: two interleaved transforms stay in registers
constexpr int kPipeFrames = 2;     // NS: frames per unit
constexpr int kPipeStages = 2;     // ring slots per warp

// Measured alternatives on B200 (same 22.8 us at the north-star shape for all of them, the kernel is bound by
// shared-memory wavefronts + FP32 pipe cycles per frame): 3 CTAs/SM with 48-register
// compact constants (rf::CompactConsts), one frame per unit at 4 CTAs/SM, a single ring slot refilled from
// rfft_streams' input_consumed hook.
// PAIR (NS == 2): the unit's two frames go through ONE 1024-point complex transform (cfft_pair.cuh: z = frame_0 +
// i frame_1, one exchange, mirror shuffle, addition-only separation); lane j then holds bins j + 32 r of both frames.
template <int LAYOUT, bool DOUBLE_INTERIOR, bool SHIFT256, int NS = kPipeFrames, int CTAS = kPipeCtasPerSm,
          bool COMPACT = false, int STAGES = kPipeStages, bool PAIR = false>
__global__ void __launch_bounds__(32 * kPipeWarps, CTAS)
stft1024_warp_kernel(const float* __restrict__ x, int64_t rows, int64_t samples, int64_t row_stride,
                     int64_t pad_left, int64_t frames, int shift, const float4* __restrict__ lane_table,
                     float* __restrict__ out, int ablate, FeatureArgs feat, const float* __restrict__ window,
                     const float2* __restrict__ tab) {
  static_assert(!PAIR || NS == 2, "the pair transform takes the two frames of a unit");
  extern __shared__ __align__(16) float smem[];   // per warp: [STAGES][span] samples, NS exchange tiles, output rows
  __shared__ __align__(8) uint64_t bars[kPipeWarps][STAGES];
  constexpr int kOutPerFrame = LAYOUT <= B2S_SPEC_CONCAT ? 2 * rf::kBins : rf::kBins;
  constexpr int kOutArea = (NS * kOutPerFrame + 8 + 3) / 4 * 4;   // NS rows + alignment slack, multiple of 4 floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int span = (NS - 1) * shift + rf::kSize;    // NS frames, `shift` apart
  const int warp_floats = STAGES * span + 2 * NS * rf::kTile1 + kOutArea;
  float* ring = smem + warp * warp_floats;
  float2* tile = reinterpret_cast<float2*>(ring + STAGES * span);
  float* obuf = ring + STAGES * span + 2 * NS * rf::kTile1;
  uint64_t* bar = bars[warp];
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) tma::mbar_init(bar + i, 1);
    tma::fence_mbar_init();
  }
  __syncwarp();
  typename std::conditional<COMPACT, rf::CompactConsts, rf::LaneConsts>::type k;
  cp::PairConsts kp;
  if (PAIR) {   // immutable plan data: may be read before the preceding kernel has finished
    kp.lane = lane;
    // the separation's factor 1/2 is part of the window; DOUBLE_INTERIOR (adjoint of the iSTFT) wants the interior
    // bins doubled: the window stays whole and the two edge bins are halved instead
#pragma unroll
    for (int p = 0; p < 32; ++p) kp.w[p] = __ldg(window + lane + 32 * p);   // pre-scaled by the launcher's choice of table
#pragma unroll
    for (int q = 0; q < 32; ++q) kp.t[cp::out_pos(q)] = __ldg(tab + 32 * q + lane);   // plan->pair_tw: [q][lane]
  } else {
    k.load(lane_table, lane);
  }
  // Launched with programmatic stream serialization: barriers and constants above overlap the tail of the
  // preceding kernel; nothing of the caller's data is touched before this wait.  Only then may a kernel launched
  // the same way behind this one (the fused loss) start -- it reads its own inputs before ITS wait and must not
  // overtake this kernel's predecessor.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int kS = LAYOUT == B2S_SPEC_INTERLEAVED ? 2 : 1;
  const int offa0 = kS * rf::bin_a(lane, 0), offa4 = kS * (rf::bin_a(lane, 4) - 256);
  const int offb0 = kS * rf::kHalf - offa0, offb4 = kS * rf::kHalf - offa4;

  // Units are enumerated as (row, index in row); a warp steps through them with a fixed stride, tracked
  // incrementally: the steady-state loop has no divisions and one 64-bit multiplication.  Unit counts and
  // sample offsets fit 32 bits (checked by the launcher).
  const int upr = (int)ceil_div(frames, NS);                       // units per row
  const int nwarps = (int)gridDim.x * kPipeWarps;
  const int step_row = nwarps / upr, step_idx = nwarps - step_row * upr;
  const int nrows_i = (int)rows, nsamples = (int)samples, pad = (int)pad_left;
  struct Pos { int row, idx; };
  auto advance = [&](Pos& p) {
    p.row += step_row;
    p.idx += step_idx;
    if (p.idx >= upr) { p.idx -= upr; ++p.row; }
  };
  unsigned parity = 0;      // bit i: phase of slot i's barrier the warp waits for next
  unsigned by_tma = 0;      // bit i: slot i is being filled by a bulk copy

  // start filling `slot` with the unit at p (TMA, or zero-filling cp.async for units that touch the zero padding
  // or are not 16-byte aligned)
  auto issue = [&](Pos p, int slot) {
    if (p.row < nrows_i) {
      const int s0 = p.idx * NS * shift - pad;
      const float* xr = x + (int64_t)p.row * row_stride;
      const bool a16 = (s0 & 3) == 0 && (reinterpret_cast<uintptr_t>(xr) & 15) == 0;
      const bool bulk = a16 && s0 >= 0 && s0 + span <= nsamples;
      by_tma = bulk ? (by_tma | (1u << slot)) : (by_tma & ~(1u << slot));
      if (bulk) {
        if (lane == 0) {
          // (no fence.proxy.async: the slot was only READ through the generic proxy, see fused.cu)
          tma::mbar_expect_tx(bar + slot, (unsigned)span * 4u);
          tma::bulk_g2s(ring + slot * span, xr + s0, (unsigned)span * 4u, bar + slot);
        }
      } else if (a16) {   // just as asynchronous as the bulk copy (cp.async.wait_group before pass 1)
        for (int c = lane; c < span / 4; c += 32) {
          const int n = s0 + 4 * c;
          const int bytes = n < 0 ? 0 : max(0, min(4, nsamples - n)) * 4;
          fft::cp_async_16(ring + slot * span + 4 * c, bytes ? xr + n : xr, bytes);
        }
      } else {
        for (int i = lane; i < span; i += 32) {
          const int n = s0 + i;
          const bool ok = n >= 0 && n < nsamples;
          fft::cp_async_4_zfill(ring + slot * span + i, ok ? xr + n : xr, ok ? 4 : 0);
        }
      }
    }
    fft::cp_async_commit();   // one (possibly empty) group per call keeps the wait_group counting uniform
  };

  Pos cur;
  {
    const int u0 = (int)blockIdx.x * kPipeWarps + warp;
    cur.row = u0 / upr;
    cur.idx = u0 - cur.row * upr;
  }
  Pos nxt = cur;
  issue(cur, 0);
  for (int it = 0; cur.row < nrows_i; ++it) {
    const int slot = it % STAGES;
    advance(nxt);
    if (STAGES > 1) issue(nxt, (it + 1) % STAGES);
    const int row = cur.row, m0 = cur.idx * NS;
    cur = nxt;
    float* buf = ring + slot * span;
    if (by_tma & (1u << slot)) {
      tma::mbar_wait(bar + slot, (parity >> slot) & 1u);
      parity ^= 1u << slot;
    } else {
      // the unit's own cp.async group is complete once at most the groups issued after it are pending
      if (STAGES > 1) fft::cp_async_wait_group<STAGES - 1>(); else fft::cp_async_wait_all();
      __syncwarp();
    }
    float2 ya[NS][8], yb[NS][8];
    float ydc[NS], ynyq[NS];
    auto next_copy = [&]() { if (STAGES == 1) issue(nxt, 0); };
    float2 pa[PAIR ? 16 : 1], pb[PAIR ? 16 : 1];   // PAIR: spectra of frame 0 / frame 1 at bins lane + 32 r
    float2 pn = make_float2(0.f, 0.f);             // PAIR, lane 0: (Y_0[512], Y_1[512]), both real
    if (PAIR) {
      float2 v[32];
      {
        const float* fa = buf + lane;
        if (SHIFT256) {   // frame 1 = frame 0 shifted by eight of the lane's strides: 40 loads instead of 64
          float xs[40];
#pragma unroll
          for (int p = 0; p < 40; ++p) xs[p] = fa[32 * p];
#pragma unroll
          for (int p = 0; p < 32; ++p) v[p] = make_float2(kp.w[p] * xs[p], kp.w[p] * xs[p + 8]);
        } else {
          const float* fb = fa + shift;
#pragma unroll
          for (int p = 0; p < 32; ++p) v[p] = make_float2(kp.w[p] * fa[32 * p], kp.w[p] * fb[32 * p]);
        }
      }
      __syncwarp();   // every lane holds its samples; the previous unit's column reads of the tile are done
      next_copy();
      cp::radix32(v);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float2 u = v[cp::out_pos(q)];
        tile[lane * cp::kPitch + q] = q == 0 ? u : rf::cmul(u, kp.t[cp::out_pos(q)]);
      }
      __syncwarp();
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = tile[l * cp::kPitch + lane];
      cp::radix32(v);
      const int partner = (32 - lane) & 31;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 z = v[cp::out_pos(r)];
        const float2 hi = v[cp::out_pos(31 - r)], lo = v[cp::out_pos((32 - r) & 31)];
        const float2 send = lane == 0 ? lo : hi;
        float2 mm;
        mm.x = __shfl_sync(0xffffffffu, send.x, partner);
        mm.y = __shfl_sync(0xffffffffu, send.y, partner);
        const float2 cm = make_float2(mm.x, -mm.y);
        const float2 sa = rf::add2(z, cm), d = rf::sub2(z, cm);
        pa[r] = sa;                          // Y_0[k] = (Z[k] + conj Z[1024 - k]) / 2
        pb[r] = make_float2(d.y, -d.x);      // Y_1[k] = (Z[k] - conj Z[1024 - k]) / 2i
      }
      const float2 zn = v[cp::out_pos(16)];
      // bin 512 (lane 0): Z[512] = (Y_0 + i Y_1) / 2 resp. Y_0 + i Y_1 with the whole window
      pn = DOUBLE_INTERIOR ? zn : make_float2(2.f * zn.x, 2.f * zn.y);
      if (DOUBLE_INTERIOR && lane == 0) { pa[0] = make_float2(0.5f * pa[0].x, 0.f); pb[0] = make_float2(0.5f * pb[0].x, 0.f); }
    } else {
      rf::rfft_streams<NS, SHIFT256 && (NS > 1), DOUBLE_INTERIOR>(buf, shift, tile, k, ya, yb, ydc, ynyq, ablate, next_copy);
    }
    if (ablate & 1) continue;   // experiments: no output at all
    // The spectrum rows of the unit are adjacent in global memory.  They are assembled in shared memory at
    // the global address's phase within 16 bytes and leave as ONE asynchronous TMA bulk store (plus at most 3
    // floats at either end from lanes): per-lane STG of rows that are only 4-byte aligned costs several LSU
    // cycles per touched line and blocks the shared-memory traffic of the whole SM behind it.
    const int nrows = min(NS, (int)frames - m0);
    constexpr bool kFeature = LAYOUT == B2S_SPEC_FEATURE;
    const bool mel = kFeature && feat.filters > 0;
    float* g = out + ((int64_t)row * frames + m0) * (mel ? feat.filters : kOutPerFrame);
    const int phase = mel ? 0 : (int)((reinterpret_cast<uintptr_t>(g) & 15) >> 2);
    if (lane == 0) tma::bulk_wait_read<0>();   // the previous unit's store has read the staging rows
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      float* o = obuf + phase + s * kOutPerFrame;
      if (PAIR) {   // bins lane + 32 r: unit-stride, conflict-free staging stores with immediates
        float* ol = o + kS * lane;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 y = s == 0 ? pa[r] : pb[r];
          if (kFeature) {
            const float va = feature_power(y, feat);
            ol[32 * r] = mel ? va : feature_log(va, feat);
          } else {
            store_bin_t<LAYOUT>(ol + kS * 32 * r, 0, y);
          }
        }
        if (lane == 0) {
          const float2 yn = make_float2(s == 0 ? pn.x : pn.y, 0.f);
          if (kFeature) {
            const float vn = feature_power(yn, feat);
            o[rf::kHalf] = mel ? vn : feature_log(vn, feat);
          } else {
            store_bin_t<LAYOUT>(o, rf::kHalf, yn);
          }
        }
        continue;
      }
      if (kFeature) {
        // |Y|^power [/ size], and without a filterbank the logarithm, per bin
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float va = feature_power(ya[s][p], feat), vb = feature_power(yb[s][p], feat);
          o[(p < 4 ? offa0 : offa4) + 64 * p] = mel ? va : feature_log(va, feat);
          o[(p < 4 ? offb0 : offb4) - 64 * p] = mel ? vb : feature_log(vb, feat);
        }
        if (lane == 0) {
          const float vd = feature_power(make_float2(ydc[s], 0.f), feat), vn = feature_power(make_float2(ynyq[s], 0.f), feat);
          o[0] = mel ? vd : feature_log(vd, feat);
          o[rf::kHalf] = mel ? vn : feature_log(vn, feat);
        }
        continue;
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        store_bin_t<LAYOUT>(o + (p < 4 ? offa0 : offa4) + kS * 64 * p, 0, ya[s][p]);
        store_bin_t<LAYOUT>(o + (p < 4 ? offb0 : offb4) - kS * 64 * p, 0, rf::conj(yb[s][p]));
      }
      if (lane == 0) {
        store_bin_t<LAYOUT>(o, 0, make_float2(ydc[s], 0.f));
        store_bin_t<LAYOUT>(o, rf::kHalf, make_float2(ynyq[s], 0.f));
      }
    }
    if (mel) {
      // filterbank: filter m = sum over its (contiguous) support of weight * |Y|^power, then the logarithm; lanes
      // own filters m = lane, lane + 32, ...; rows of `filters` floats leave with plain stores (80 floats per frame
      // against 513: the output stream shrinks 6.4 x)
      __syncwarp();
      for (int s = 0; s < nrows; ++s) {
        const float* o = obuf + s * kOutPerFrame;
        for (int m = lane; m < feat.filters; m += 32) {
          const int lo = __ldg(feat.lo + m), len = __ldg(feat.len + m);
          const float* w = feat.weights + __ldg(feat.woff + m);
          float acc = 0.f;
          for (int j = 0; j < len; ++j) acc = fmaf(__ldg(w + j), o[lo + j], acc);
          g[(int64_t)s * feat.filters + m] = feature_log(acc, feat);
        }
      }
      continue;   // the next unit's bulk_wait_read / __syncwarp orders these reads before its writes
    }
    tma::fence_proxy_async();   // the rows were written through the generic proxy
    __syncwarp();
    const int n = nrows * kOutPerFrame;
    const int head = (4 - phase) & 3, mid = (n - head) & ~3, tail = n - head - mid;
    if (lane == 0) {
      tma::bulk_s2g(g + head, obuf + phase + head, (unsigned)mid * 4u);
      tma::bulk_commit();
    }
    if (lane < head) g[lane] = obuf[phase + lane];
    if (lane < tail) g[head + mid + lane] = obuf[phase + head + mid + lane];
  }
  if (lane == 0) tma::bulk_wait<0>();   // shared memory must outlive the last store
}

// ------------------------------------------------------------------------------------------- generic forward
// One CTA per frame: the windowed frame is staged in shared memory, thread f accumulates bin f with
// the twiddle index (f k mod size) advanced incrementally.
__global__ void __launch_bounds__(256)
dft_forward_kernel(const float* __restrict__ x, int64_t rows, int64_t samples, int64_t row_stride,
                   int64_t pad_left, int64_t frames, int size, int shift, int wlen,
                   const float* __restrict__ win, const float2* __restrict__ twtab, int layout,
                   float interior_scale, float* __restrict__ out) {
  extern __shared__ float frame[];
  const int bins = size / 2 + 1;
  for (int64_t fr = blockIdx.x; fr < rows * frames; fr += gridDim.x) {
    const int64_t row = fr / frames, m = fr - row * frames;
    const float* xr = x + row * row_stride;
    const int64_t s0 = m * shift - pad_left;
    __syncthreads();
    for (int k = threadIdx.x; k < wlen; k += blockDim.x) {
      const int64_t i = s0 + k;
      frame[k] = (i >= 0 && i < samples) ? __ldg(xr + i) * win[k] : 0.f;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < bins; f += blockDim.x) {
      float re = 0.f, im = 0.f;
      int idx = 0;
      for (int k = 0; k < wlen; ++k) {
        const float2 w = __ldg(twtab + idx);
        const float v = frame[k];
        re = fmaf(v, w.x, re);
        im = fmaf(v, w.y, im);
        idx += f;
        if (idx >= size) idx -= size;
      }
      const float sc = (f == 0 || f == size / 2) ? 1.f : interior_scale;
      store_bin(out, fr, f, make_float2(re * sc, im * sc), layout, bins);
    }
  }
}

// ------------------------------------------------------------------------------------------- fast inverse
// CTA = (row, chunk of `hops` output hops).  Phase 1: the frames overlapping the chunk are inverse
// transformed two at a time per warp (rf::irfft_streams: the forward passes run on conj Z), windowed and
// parked in shared memory (frame slot == spectrum landing area == exchange tile).  Phase 2: every output
// sample sums its <= ceil(wlen/shift) contributions in increasing frame order -- deterministic, no atomics.
constexpr int kInvWarps = 4;
// NS = 1: 16 slots (72 KB) and <= 168 registers -> 3 CTAs per SM; NS = 2: 24 slots (108 KB), 2 CTAs per SM
__host__ __device__ constexpr int inv_slots(int ns) { return ns == 1 ? 16 : 24; }

// stage the 513 bins of frame `fr` (global frame index) as interleaved float2 into `slot`
template <int LAYOUT>
__device__ __forceinline__ void stage_spectrum(float2* slot, const float* __restrict__ spec, int64_t fr, int lane) {
  if (LAYOUT == B2S_SPEC_INTERLEAVED) {
    // bin lane + 32 i lands at inv_bin_pos(lane) + 16 i: one base per side and immediates (16 copies + bin 512 from lane 0)
    const float2* src = reinterpret_cast<const float2*>(spec) + fr * rf::kBins + lane;
    float2* dst = slot + rf::inv_bin_pos(lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) fft::cp_async_8(dst + 16 * i, src + 32 * i);
    if (lane == 0) fft::cp_async_8(dst + 256, src + 512);
  } else {   // 'concat': a row of real parts, then a row of imaginary parts
    const float* re = spec + fr * 2 * rf::kBins;
    float* dst = reinterpret_cast<float*>(slot);
    for (int c = lane; c < rf::kBins; c += 32) {
      fft::cp_async_4(dst + 2 * rf::inv_bin_pos(c), re + c);
      fft::cp_async_4(dst + 2 * rf::inv_bin_pos(c) + 1, re + rf::kBins + c);
    }
  }
}

template <int LAYOUT, int NS>
__global__ void __launch_bounds__(32 * kInvWarps, NS == 1 ? 3 : 2)
istft1024_kernel(const float* __restrict__ spec, int64_t rows, int64_t frames, int shift,
                 int wlen, int overlap /*ceil(wlen/shift)*/, int hops, int64_t chunks,
                 int64_t crop_left, int64_t samples_out, const float4* __restrict__ lane_table,
                 float interior_in_scale, float* __restrict__ out) {
  extern __shared__ float2 slots[];  // [inv_slots(NS)][rf::kTile1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  rf::InvLaneConsts k;
  k.load(lane_table, lane, interior_in_scale);
  // 16-byte overlap-add: whole float4 groups stay inside one frame's window and one output alignment class
  const bool vec4 = (shift & 3) == 0 && (wlen & 3) == 0 && (crop_left & 3) == 0 && (samples_out & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int64_t job = blockIdx.x; job < rows * chunks; job += gridDim.x) {
    const int64_t row = job / chunks, chunk = job - row * chunks;
    const int64_t h0 = chunk * hops;                       // first hop (padded sample h0*shift)
    const int64_t m_first = max((int64_t)0, h0 - overlap + 1);
    const int64_t m_last = min(frames - 1, h0 + hops - 1);  // inclusive
    __syncthreads();                                        // the previous job's overlap-add is done
    // this warp's frame groups (m .. m + NS - 1), m = m_first + NS (warp + 4 i): all spectra start travelling
    // now, one cp.async group per frame group
    int ngroups = 0;
    for (int64_t m = m_first + NS * warp; m <= m_last; m += NS * kInvWarps, ++ngroups) {
      float2* slot = slots + (m - m_first) * rf::kTile1;
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (m + s <= m_last) stage_spectrum<LAYOUT>(slot + s * rf::kTile1, spec, row * frames + m + s, lane);
      fft::cp_async_commit();
    }
    int grp = 0;
    for (int64_t m = m_first + NS * warp; m <= m_last; m += NS * kInvWarps, ++grp) {
      float2* tile = slots + (m - m_first) * rf::kTile1;
      // groups complete in order: group `grp` has landed once at most ngroups - 1 - grp groups are pending
      switch (ngroups - 1 - grp) {
        case 0: fft::cp_async_wait_group<0>(); break;
        case 1: fft::cp_async_wait_group<1>(); break;
        case 2: fft::cp_async_wait_group<2>(); break;
        default: fft::cp_async_wait_group<3>(); break;
      }
      __syncwarp();
      float2 a[NS][8], b[NS][8];
      rf::irfft_streams<NS>(tile, k, a, b);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float2* fr = tile + s * rf::kTile1;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          fr[lane + 64 * p] = a[s][p];
          fr[rf::inv_pos_b(lane, p)] = b[s][p];
        }
      }
    }
    __syncthreads();
    // overlap-add: output sample p = (h0 + hq) * shift + i sums frame m = h0 + hq - j at index i + j * shift,
    // j = overlap-1 .. 0 (increasing frame order); all index arithmetic relative to the chunk, in 32 bits
    const float* fbuf = reinterpret_cast<const float*>(slots);
    const int span = hops * shift;
    const int rel0 = (int)(h0 - m_first);            // slot of frame h0 (0 .. overlap-1)
    const int last_rel = (int)(m_last - m_first);    // last valid slot
    const int64_t n0 = h0 * shift - crop_left;       // output index of the chunk's first sample
    float* orow = out + row * samples_out;
    if (vec4) {
      // four consecutive samples per thread: one LDS.128 per contributing frame, one 16-byte store
      for (int q4 = threadIdx.x; q4 < (span >> 2); q4 += blockDim.x) {
        const int q = q4 << 2;
        const int64_t n = n0 + q;
        if (n + 3 < 0 || n >= samples_out) continue;
        const int hq = q / shift, i = q - hq * shift;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = overlap - 1; j >= 0; --j) {
          const int rel = rel0 + hq - j, idx = i + j * shift;
          if (rel >= 0 && rel <= last_rel && idx < wlen) {
            const float4 v = *reinterpret_cast<const float4*>(fbuf + rel * (2 * rf::kTile1) + idx);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
        if (n >= 0 && n + 3 < samples_out) {
          *reinterpret_cast<float4*>(orow + n) = acc;
        } else {
          const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
          for (int c = 0; c < 4; ++c)
            if (n + c >= 0 && n + c < samples_out) orow[n + c] = a4[c];
        }
      }
    } else {
      for (int q = threadIdx.x; q < span; q += blockDim.x) {
        const int64_t n = n0 + q;
        if (n < 0 || n >= samples_out) continue;
        const int hq = q / shift, i = q - hq * shift;
        float acc = 0.f;
        for (int j = overlap - 1; j >= 0; --j) {
          const int rel = rel0 + hq - j, idx = i + j * shift;
          if (rel >= 0 && rel <= last_rel && idx < wlen) acc += fbuf[rel * (2 * rf::kTile1) + idx];
        }
        orow[n] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- ring inverse
// Halo-free overlap-add for shift 256 (four hops per frame): every row is cut into chunks of consecutive frames
// (at least four), one chunk = one unit of a persistent grid of independent WARPS.  A warp walks its chunk NS
// frames at a time: the spectra of the next step travel into the other half of its double buffer (cp.async)
// while rf::irfft_streams transforms the current ones in place.  The overlap-add never touches shared memory: a
// lane's windowed samples fall on the SAME four 8-byte positions of whichever hop they belong to (a side: sample
// pairs lane and 64 + lane of the hop, b side: 64 - lane and 128 - lane), so the sums of the hops in flight live
// in registers (a shift register of 3 + NS hops x 4 float2), contributions arrive in increasing frame order (the
// sums equal the sequential overlap-add), and a completed hop leaves as four warp-wide contiguous 256-byte stores.
// No block barrier, no frame is transformed twice.  The three hops on either side of a chunk boundary receive
// contributions from both neighbours: each side writes its partial sums to the workspace and takes a ticket;
// whoever comes second adds the two (a + b is commutative: the result does not depend on who that is) and writes
// the output -- nobody waits.
// workspace = [kMaxTickets ints, zero between calls][boundaries][2 sides][3 hops][4][32 lanes] float2.
constexpr int kRingHop = 256, kRingWarps = 4;
__host__ __device__ constexpr int ring_ctas(int ns) { return ns == 1 ? 3 : 2; }
__host__ __device__ constexpr int ring_warp_floats(int ns) { return 2 * ns * 2 * rf::kTile1; }   // double-buffered spectra
constexpr int kBoundaryFloats = 2 * 3 * kRingHop;
struct RingGeometry { int64_t chunks_per_row, units, boundaries; int grid; };
RingGeometry ring_geometry(int64_t rows, int64_t frames, int ns) {
  RingGeometry g;
  const int64_t warps = (int64_t)kNumSMs * ring_ctas(ns) * kRingWarps;
  g.chunks_per_row = std::max<int64_t>(1, std::min<int64_t>(warps / std::max<int64_t>(1, rows), frames / 4));
  g.units = rows * g.chunks_per_row;
  g.boundaries = rows * (g.chunks_per_row - 1);
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(g.units, kRingWarps), (int64_t)kNumSMs * ring_ctas(ns)));
  return g;
}

template <int LAYOUT, int NS>
__global__ void __launch_bounds__(32 * kRingWarps, ring_ctas(NS))
istft_ring_kernel(const float* __restrict__ spec, int64_t rows, int64_t frames, int64_t chunks_per_row,
                  int64_t crop_left, int64_t samples_out, const float4* __restrict__ lane_table,
                  float interior_in_scale, float* __restrict__ out, int* __restrict__ tickets,
                  float* __restrict__ partials) {
  extern __shared__ __align__(16) float ring_smem[];
  __shared__ __align__(8) uint64_t ring_bars[kRingWarps][2];
  // interleaved complex rows (4104 contiguous bytes, 8-byte aligned) travel by ONE bulk copy each and are read in bin
  // order; 'concat' rows are staged bin by bin with cp.async
  constexpr bool kBulk = LAYOUT == B2S_SPEC_INTERLEAVED;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* bufs = reinterpret_cast<float2*>(ring_smem + warp * ring_warp_floats(NS));   // [2][NS][kTile1]
  if (kBulk && lane == 0) {
    tma::mbar_init(&ring_bars[warp][0], 1);
    tma::mbar_init(&ring_bars[warp][1], 1);
    tma::fence_mbar_init();
  }
  __syncwarp();
  unsigned phase_bits = 0;   // bit `which`: parity the next wait on that buffer's barrier expects
  unsigned off_bits = 0;     // bit which * NS + s: float2 offset (0 or 1) of bin 0 inside its landing region
  rf::InvLaneConsts k;
  k.load(lane_table, lane, interior_in_scale);
  const int64_t units = rows * chunks_per_row;
  const int64_t nwarps = (int64_t)gridDim.x * kRingWarps;
  const int64_t total_hops = (crop_left + samples_out + kRingHop - 1) / kRingHop;   // hops that can receive output
  // sample-pair positions of a lane inside a hop: v[0] lane, v[1] 64 + lane, v[2] 64 - lane (lane 0: 32), v[3] 64 + that
  const int qb = lane ? 64 - lane : 32;
  auto store_hop = [&](float* orow, int64_t h, const float2 (&v)[4]) {
    if (h >= total_hops) return;
    const int64_t n0 = h * kRingHop - crop_left;
    if (n0 >= 0 && n0 + kRingHop <= samples_out) {   // warp-uniform: the whole hop lies inside the output row
      float2* o = reinterpret_cast<float2*>(orow + n0);
      o[lane] = v[0];
      o[64 + lane] = v[1];
      o[qb] = v[2];
      o[64 + qb] = v[3];
      return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t n = n0 + 2 * ((q & 1) * 64 + (q < 2 ? lane : qb));
      if (n >= 0 && n < samples_out) orow[n] = v[q].x;
      if (n + 1 >= 0 && n + 1 < samples_out) orow[n + 1] = v[q].y;
    }
  };
  auto park_hop = [&](int64_t bid, int side, int j, const float2 (&v)[4]) {
    float2* dst = reinterpret_cast<float2*>(partials + bid * kBoundaryFloats) + ((side * 3 + j) * 4) * 32 + lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q * 32] = v[q];
  };
  // Publish the parked boundary hops of a chunk -- its leading ones (boundary bid_lead, if any) and its trailing ones
  // (bid_trail, if any) -- with ONE fence and one round of tickets (lanes 0 and 1), then combine wherever the
  // neighbour was first.
  auto publish = [&](bool lead, bool trail, int64_t bid_lead, int64_t bid_trail, int64_t row, int64_t f0, int64_t f1) {
    if (!lead && !trail) return;
    __threadfence();
    __syncwarp();
    int old = 0;
    if (lane == 0 && lead) old = atomicAdd(tickets + bid_lead, 1);
    if (lane == 1 && trail) old = atomicAdd(tickets + bid_trail, 1);
    const int old_lead = __shfl_sync(0xffffffffu, old, 0), old_trail = __shfl_sync(0xffffffffu, old, 1);
    const bool second_lead = lead && old_lead == 1, second_trail = trail && old_trail == 1;
    if (!second_lead && !second_trail) return;   // warp-uniform
    __threadfence();
    float* orow = out + row * samples_out;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (!(w == 0 ? second_lead : second_trail)) continue;
      const int64_t bid = w == 0 ? bid_lead : bid_trail, h_first = w == 0 ? f0 : f1;
      const float2* e = reinterpret_cast<const float2*>(partials + bid * kBoundaryFloats) + lane;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float2 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a = __ldcg(e + (j * 4 + q) * 32), b = __ldcg(e + ((3 + j) * 4 + q) * 32);
          v[q] = make_float2(a.x + b.x, a.y + b.y);
        }
        store_hop(orow, h_first + j, v);
      }
      if (lane == 0) tickets[bid] = 0;
    }
  };
  for (int64_t u = (int64_t)blockIdx.x * kRingWarps + warp; u < units; u += nwarps) {
    const int64_t row = u / chunks_per_row, c = u - row * chunks_per_row;
    const int64_t f0 = frames * c / chunks_per_row, f1 = frames * (c + 1) / chunks_per_row;
    float* orow = out + row * samples_out;
    const bool lead_partial = f0 > 0, trail_partial = f1 < frames;
    const int64_t bid_lead = row * (chunks_per_row - 1) + (c - 1), bid_trail = bid_lead + 1;
    auto stage_step = [&](int which, int64_t m) {
      if (kBulk) {
        unsigned bytes[NS], total = 0;
        uintptr_t from[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const uintptr_t addr = reinterpret_cast<uintptr_t>(spec) + (uintptr_t)(row * frames + m + s) * (rf::kBins * 8);
          off_bits = (off_bits & ~(1u << (which * NS + s))) | ((unsigned)((addr >> 3) & 1) << (which * NS + s));
          from[s] = addr & ~(uintptr_t)15;
          bytes[s] = m + s < f1 ? (unsigned)(((addr & 15) + rf::kBins * 8 + 15) & ~15u) : 0u;
          total += bytes[s];
        }
        if (lane == 0) {
          // the region was the exchange buffer of the transform before last (generic-proxy stores and loads, all
          // finished before the __syncwarp() that ended it): order them before the copy's async-proxy writes
          tma::fence_proxy_async();
          tma::mbar_expect_tx(&ring_bars[warp][which], total);
#pragma unroll
          for (int s = 0; s < NS; ++s)
            if (bytes[s]) tma::bulk_g2s(bufs + (which * NS + s) * rf::kTile1, reinterpret_cast<const void*>(from[s]), bytes[s], &ring_bars[warp][which]);
        }
        return;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (m + s < f1) stage_spectrum<LAYOUT>(bufs + (which * NS + s) * rf::kTile1, spec, row * frames + m + s, lane);
      fft::cp_async_commit();
    };
    stage_step(0, f0);
    float2 acc[3 + NS][4];   // hop m + t of the current step; slots 0..2 carry sums of earlier frames
#pragma unroll
    for (int t = 0; t < 3 + NS; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[t][q] = make_float2(0.f, 0.f);
    int which = 0;
    for (int64_t m = f0; m < f1; m += NS, which ^= 1) {
      const bool more = m + NS < f1;
      if (more) stage_step(which ^ 1, m + NS);
      if (kBulk) {
        tma::mbar_wait(&ring_bars[warp][which], (phase_bits >> which) & 1u);
        phase_bits ^= 1u << which;
      } else {
        if (more) fft::cp_async_wait_group<1>(); else fft::cp_async_wait_group<0>();
        __syncwarp();
      }
      float2 a[NS][8], b[NS][8];
      int spec_off[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) spec_off[s] = (int)((off_bits >> (which * NS + s)) & 1u);
      if (NS == 2 && m + 1 >= f1) {   // warp-uniform: an odd chunk's last frame is transformed alone (60 % of a pair's time)
        float2 a1[1][8], b1[1][8];
        rf::irfft_streams<1, kBulk>(bufs + which * NS * rf::kTile1, k, a1, b1, spec_off);
#pragma unroll
        for (int p = 0; p < 8; ++p) { a[0][p] = a1[0][p]; b[0][p] = b1[0][p]; }
      } else {
        rf::irfft_streams<NS, kBulk>(bufs + which * NS * rf::kTile1, k, a, b, spec_off);
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (m + s < f1) {   // warp-uniform
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // hop j of the frame: a-side pairs 2 j, 2 j + 1; b-side slots whose r = (p + 7) & 7 is 2 j, 2 j + 1
            const float2 c0 = a[s][2 * j], c1 = a[s][2 * j + 1], c2 = b[s][(2 * j + 1) & 7], c3 = b[s][(2 * j + 2) & 7];
            float2 (&dst)[4] = acc[s + j];
            if (j == 3) { dst[0] = c0; dst[1] = c1; dst[2] = c2; dst[3] = c3; }   // first contribution to that hop
            else { dst[0] = rf::add2(dst[0], c0); dst[1] = rf::add2(dst[1], c1); dst[2] = rf::add2(dst[2], c2); dst[3] = rf::add2(dst[3], c3); }
          }
        }
      }
      // hops m .. m + NS - 1 are complete as far as this chunk is concerned
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int64_t h = m + s;
        if (h < f1) {
          if (lead_partial && h < f0 + 3) {
            park_hop(bid_lead, 1, (int)(h - f0), acc[s]);   // published together with the trailing hops
          } else {
            store_hop(orow, h, acc[s]);
          }
        }
      }
      if (m + NS <= f1) {   // shift the register ring by NS hops (a short last step keeps its slots: see below)
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[t][q] = acc[t + NS][q];
      }
    }
    // the three hops after the last frame: slots 0..2 after a full last step, slots 1..3 after a short one (NS = 2, odd count)
    const bool short_last = NS == 2 && ((f1 - f0) & 1);
    float2 tail[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) tail[j][q] = short_last ? acc[(j + 1) % (3 + NS)][q] : acc[j][q];
    if (trail_partial) {
#pragma unroll
      for (int j = 0; j < 3; ++j) park_hop(bid_trail, 0, j, tail[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) store_hop(orow, f1 + j, tail[j]);
      // positions no frame reaches (a requested length beyond the frames' support) are zeros
      const float2 z[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      for (int64_t h = f1 + 3; h < total_hops; ++h) store_hop(orow, h, z);
    }
    publish(lead_partial, trail_partial, bid_lead, bid_trail, row, f0, f1);
  }
}

// ------------------------------------------------------------------------------------------- pair ring inverse
// The ring inverse with the pair transform run backwards (interleaved complex rows, shift 256): the spectra A, B of
// the two frames of a step form ONE Hermitian-extended 1024-point complex spectrum Z = A + i B (Z[1024 - k] =
// conj A[k] + i conj B[k]: the mirrored half is read from the same landed rows, no exchange), whose inverse transform
// z = a + i b carries frame m in its real and frame m + 1 in its imaginary part.  z = conj(FFT(conj Z)) runs through
// the forward machinery of cfft_pair.cuh (radix-32, one exchange, radix-32); lane j ends up with samples j + 32 q of
// both frames -- eight per hop at the same positions for every frame, so the overlap-add is a register shift register
// again (5 hops x 8 floats) and a completed hop leaves as eight warp-wide contiguous 128-byte stores.  Chunks, chunk
// boundaries (ticketed workspace) and the summation order are those of istft_ring_kernel.
__global__ void __launch_bounds__(32 * kRingWarps, 2)
istft_pair_ring_kernel(const float* __restrict__ spec, int64_t rows, int64_t frames, int64_t chunks_per_row,
                       int64_t crop_left, int64_t samples_out, const float* __restrict__ window,
                       const float2* __restrict__ tab, float scale /* 1: iSTFT, 1/2: adjoint of the STFT */,
                       float* __restrict__ out, int* __restrict__ tickets, float* __restrict__ partials) {
  extern __shared__ __align__(16) float ring_smem[];
  __shared__ __align__(8) uint64_t ring_bars[kRingWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* bufs = reinterpret_cast<float2*>(ring_smem + warp * ring_warp_floats(2));   // [2][2][kTile1]
  if (lane == 0) {
    tma::mbar_init(&ring_bars[warp][0], 1);
    tma::mbar_init(&ring_bars[warp][1], 1);
    tma::fence_mbar_init();
  }
  __syncwarp();
  cp::PairConsts kp;
  kp.lane = lane;
#pragma unroll
  for (int p = 0; p < 32; ++p) kp.w[p] = scale * __ldg(window + lane + 32 * p);
#pragma unroll
  for (int q = 0; q < 32; ++q) kp.t[cp::out_pos(q)] = __ldg(tab + 32 * q + lane);
  const float dc_fix = 1.f / scale;   // DC and Nyquist are not scaled
  const bool first = lane == 0;
  unsigned phase_bits = 0, off_bits = 0;
  const int64_t units = rows * chunks_per_row;
  const int64_t nwarps = (int64_t)gridDim.x * kRingWarps;
  const int64_t total_hops = (crop_left + samples_out + kRingHop - 1) / kRingHop;
  // a lane's samples of a hop: lane + 32 i, i = 0..7
  auto store_hop = [&](float* orow, int64_t h, const float (&v)[8]) {
    if (h >= total_hops) return;
    const int64_t n0 = h * kRingHop - crop_left + lane;
    if (n0 - lane >= 0 && n0 - lane + kRingHop <= samples_out) {   // warp-uniform: the whole hop lies inside the row
#pragma unroll
      for (int i = 0; i < 8; ++i) orow[n0 + 32 * i] = v[i];
      return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t n = n0 + 32 * i;
      if (n >= 0 && n < samples_out) orow[n] = v[i];
    }
  };
  auto park_hop = [&](int64_t bid, int side, int j, const float (&v)[8]) {
    float* dst = partials + bid * kBoundaryFloats + ((side * 3 + j) * 8) * 32 + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i * 32] = v[i];
  };
  auto publish = [&](bool lead, bool trail, int64_t bid_lead, int64_t bid_trail, int64_t row, int64_t f0, int64_t f1) {
    if (!lead && !trail) return;
    __threadfence();
    __syncwarp();
    int old = 0;
    if (lane == 0 && lead) old = atomicAdd(tickets + bid_lead, 1);
    if (lane == 1 && trail) old = atomicAdd(tickets + bid_trail, 1);
    const int old_lead = __shfl_sync(0xffffffffu, old, 0), old_trail = __shfl_sync(0xffffffffu, old, 1);
    const bool second_lead = lead && old_lead == 1, second_trail = trail && old_trail == 1;
    if (!second_lead && !second_trail) return;   // warp-uniform
    __threadfence();
    float* orow = out + row * samples_out;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      if (!(w == 0 ? second_lead : second_trail)) continue;
      const int64_t bid = w == 0 ? bid_lead : bid_trail, h_first = w == 0 ? f0 : f1;
      const float* e = partials + bid * kBoundaryFloats + lane;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldcg(e + (j * 8 + i) * 32) + __ldcg(e + ((3 + j) * 8 + i) * 32);
        store_hop(orow, h_first + j, v);
      }
      if (lane == 0) tickets[bid] = 0;
    }
  };
  for (int64_t u = (int64_t)blockIdx.x * kRingWarps + warp; u < units; u += nwarps) {
    const int64_t row = u / chunks_per_row, c = u - row * chunks_per_row;
    const int64_t f0 = frames * c / chunks_per_row, f1 = frames * (c + 1) / chunks_per_row;
    float* orow = out + row * samples_out;
    const bool lead_partial = f0 > 0, trail_partial = f1 < frames;
    const int64_t bid_lead = row * (chunks_per_row - 1) + (c - 1), bid_trail = bid_lead + 1;
    auto stage_step = [&](int which, int64_t m) {
      unsigned bytes[2], total = 0;
      uintptr_t from[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(spec) + (uintptr_t)(row * frames + m + s) * (rf::kBins * 8);
        off_bits = (off_bits & ~(1u << (which * 2 + s))) | ((unsigned)((addr >> 3) & 1) << (which * 2 + s));
        from[s] = addr & ~(uintptr_t)15;
        bytes[s] = m + s < f1 ? (unsigned)(((addr & 15) + rf::kBins * 8 + 15) & ~15u) : 0u;
        total += bytes[s];
      }
      if (lane == 0) {
        tma::fence_proxy_async();   // the region was this warp's transpose tile (generic-proxy stores and loads)
        tma::mbar_expect_tx(&ring_bars[warp][which], total);
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (bytes[s]) tma::bulk_g2s(bufs + (which * 2 + s) * rf::kTile1, reinterpret_cast<const void*>(from[s]), bytes[s], &ring_bars[warp][which]);
      }
    };
    stage_step(0, f0);
    float acc[5][8];   // hop m + t of the current step; slots 0..2 carry sums of earlier frames
#pragma unroll
    for (int t = 0; t < 5; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    int which = 0;
    for (int64_t m = f0; m < f1; m += 2, which ^= 1) {
      const bool more = m + 2 < f1, two = m + 1 < f1;
      if (more) stage_step(which ^ 1, m + 2);
      tma::mbar_wait(&ring_bars[warp][which], (phase_bits >> which) & 1u);
      phase_bits ^= 1u << which;
      const float2* ra = bufs + (which * 2) * rf::kTile1 + ((off_bits >> (which * 2)) & 1u);
      const float2* rb = bufs + (which * 2 + 1) * rf::kTile1 + ((off_bits >> (which * 2 + 1)) & 1u);
      // v[r] = conj Z[lane + 32 r], Z = A + i B Hermitian-extended: bins up to 512 directly, the rest from the mirror bin
      float2 v[32];
      const int kmir = rf::kHalf - lane;   // mirror of bin lane + 32 r (r >= 16) is kmir - 32 (r - 16)
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int k = r < 16 ? lane + 32 * r : kmir - 32 * (r - 16);
        float2 A = ra[k];
        float2 B = two ? rb[k] : make_float2(0.f, 0.f);
        if ((r == 0 || r == 16) && first) {   // DC / Nyquist: real, not scaled
          A = make_float2(A.x * dc_fix, 0.f);
          B = make_float2(B.x * dc_fix, 0.f);
        }
        v[r] = r < 16 ? make_float2(A.x - B.y, -A.y - B.x) : make_float2(A.x + B.y, A.y - B.x);
      }
      __syncwarp();   // every lane holds its bins: the landing regions become the transpose tile
      float2* tile = bufs + (which * 2) * rf::kTile1;   // 2 * kTile1 >= 32 * kPitch float2
      cp::radix32(v);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float2 w = v[cp::out_pos(q)];
        tile[lane * cp::kPitch + q] = q == 0 ? w : rf::cmul(w, kp.t[cp::out_pos(q)]);
      }
      __syncwarp();
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = tile[l * cp::kPitch + lane];
      __syncwarp();   // the tile's reads are done: the regions may receive the step after next
      cp::radix32(v);
      // v[q] = conj z[lane + 32 q]: frame m = Re z, frame m + 1 = Im z = -Im v; window, then the register overlap-add
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float xa = v[cp::out_pos(q)].x * kp.w[q];
        float (&dst)[8] = acc[q / 8];
        if (q / 8 == 3) dst[q % 8] = xa; else dst[q % 8] += xa;   // a frame's last hop: first contribution
      }
      if (two) {   // warp-uniform
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float xb = -v[cp::out_pos(q)].y * kp.w[q];
          float (&dst)[8] = acc[1 + q / 8];
          if (q / 8 == 3) dst[q % 8] = xb; else dst[q % 8] += xb;
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int64_t h = m + s;
        if (h < f1) {
          if (lead_partial && h < f0 + 3) park_hop(bid_lead, 1, (int)(h - f0), acc[s]);
          else store_hop(orow, h, acc[s]);
        }
      }
      if (two) {   // shift the register ring by two hops (a short last step keeps its slots)
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[t][i] = acc[t + 2][i];
      }
    }
    const bool short_last = (f1 - f0) & 1;
    float tail[3][8];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) tail[j][i] = short_last ? acc[j + 1][i] : acc[j][i];
    if (trail_partial) {
#pragma unroll
      for (int j = 0; j < 3; ++j) park_hop(bid_trail, 0, j, tail[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) store_hop(orow, f1 + j, tail[j]);
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int64_t h = f1 + 3; h < total_hops; ++h) store_hop(orow, h, z);
    }
    publish(lead_partial, trail_partial, bid_lead, bid_trail, row, f0, f1);
  }
}

// ------------------------------------------------------------------------------------------- generic inverse
// Kernel A: one CTA per frame writes win[k] * (G0 + (-1)^k G_{N/2} + c sum_{0<f<N/2} Re(G_f e^{+i theta}))
// to scratch[row, m, wlen].  Kernel B gathers the overlap-add.
__global__ void __launch_bounds__(256)
dft_inverse_frames_kernel(const float* __restrict__ spec, int64_t total_frames, int layout, int size,
                          int wlen, const float* __restrict__ win, const float2* __restrict__ twtab,
                          float interior_scale, float* __restrict__ scratch) {
  extern __shared__ float2 bins_sm[];
  const int bins = size / 2 + 1;
  for (int64_t fr = blockIdx.x; fr < total_frames; fr += gridDim.x) {
    __syncthreads();
    for (int f = threadIdx.x; f < bins; f += blockDim.x) bins_sm[f] = load_bin(spec, fr, f, layout, bins);
    __syncthreads();
    for (int k = threadIdx.x; k < wlen; k += blockDim.x) {
      float acc = 0.f;
      int idx = k;  // (f k) mod size for f = 1
      for (int f = 1; f < size / 2; ++f) {
        const float2 w = __ldg(twtab + idx);  // (cos, -sin)
        const float2 g = bins_sm[f];
        acc = fmaf(g.x, w.x, acc);
        acc = fmaf(g.y, w.y, acc);             // Re(g e^{+i theta}) = g.x cos - g.y sin
        idx += k;
        if (idx >= size) idx -= size;
      }
      const float edge = bins_sm[0].x + ((k & 1) ? -bins_sm[size / 2].x : bins_sm[size / 2].x);
      scratch[fr * wlen + k] = win[k] * fmaf(interior_scale, acc, edge);
    }
  }
}

__global__ void __launch_bounds__(256)
overlap_add_gather_kernel(const float* __restrict__ scratch, int64_t rows, int64_t frames, int shift,
                          int wlen, int64_t crop_left, int64_t samples_out, float* __restrict__ out) {
  const int64_t total = rows * samples_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / samples_out, n = i - row * samples_out;
    const int64_t p = n + crop_left;
    int64_t lo = (p - wlen + 1 <= 0) ? 0 : (p - wlen + shift) / shift;
    const int64_t hi = min(p / shift, frames - 1);
    float acc = 0.f;
    for (int64_t m = lo; m <= hi; ++m) acc += scratch[(row * frames + m) * wlen + (p - m * shift)];
    out[i] = acc;
  }
}

int fading_extra(int wlen, int shift, int fading) {
  if (fading == 0) return 0;
  return (fading == 2 ? 1 : 2) * (wlen - shift);
}

int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

int check_plan(const b2s_stft_plan* plan) {
  B2S_REQUIRE(plan != nullptr, "stft plan is NULL");
  return B2S_OK;
}

bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

// launch of one instantiation of the warp-pipeline kernel
struct PipeArgs {
  const float* x; int64_t rows, samples, row_stride, pad_left, frames; int shift; const float4* table; float* out;
  int ablate, layout, device; cudaStream_t stream; FeatureArgs feat;
  const float* window; const float2* tab;   // the pair transform's constants are built from these (plan tables)
};
template <int L, bool D, bool S, int NS = kPipeFrames, int CTAS = kPipeCtasPerSm, bool COMPACT = false,
          int STAGES = kPipeStages, bool PAIR = false>
int launch_pipe(const PipeArgs& a) {
  const int64_t units = a.rows * ceil_div(a.frames, NS);
  // B2S_FWD_SMS (tuning aid): fewer SMs for the front-end, e.g. next to a fused kernel restricted by B2S_FUSED_SMS
  static const int sms = [] { const char* e = getenv("B2S_FWD_SMS"); const int v = e ? atoi(e) : 0; return v > 0 && v < kNumSMs ? v : kNumSMs; }();
  const int grid = (int)std::min<int64_t>(ceil_div(units, kPipeWarps), (int64_t)sms * CTAS);
  const int span = (NS - 1) * a.shift + fft::kSize;
  const int out_area = (NS * (L <= B2S_SPEC_CONCAT ? 2 * fft::kBins : fft::kBins) + 8 + 3) / 4 * 4;
  const size_t smem = kPipeWarps * (sizeof(float) * (STAGES * span + out_area) + sizeof(float2) * NS * rf::kTile1);
  static const bool use_pdl = [] { const char* e = getenv("B2S_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * kPipeWarps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  auto kernel = stft1024_warp_kernel<L, D, S, NS, CTAS, COMPACT, STAGES, PAIR>;
  static bool configured[64] = {};
  if (!configured[a.device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[a.device & 63] = true;
  }
  B2S_CUDA(cudaLaunchKernelEx(&cfg, kernel, a.x, a.rows, a.samples, a.row_stride, a.pad_left, a.frames, a.shift,
                              a.table, a.out, a.ablate, a.feat, a.window, a.tab));
  B2S_LAUNCH_CHECK("stft1024_warp_kernel");
  return B2S_OK;
}

// forward-type launch shared by b2s_stft_forward and b2s_istft_backward
int launch_forward(const b2s_stft_plan* plan, const float* x, int64_t rows, int64_t samples,
                   int64_t row_stride, int64_t pad_left, int64_t frames, int layout, const float* win,
                   float interior_scale, float* out, cudaStream_t stream, const FeatureArgs* feat = nullptr) {
  const int64_t total = rows * frames;
  if (total == 0) return B2S_OK;
  if (plan->fast && plan->wlen == fft::kSize && plan->shift % 4 == 0 && plan->shift <= fft::kSize) {
    // warp pipelines (TMA-fed); any row alignment (units that are not 16-byte aligned are filled with cp.async)
    static const int ablate = [] { const char* e = getenv("B2S_ABLATE"); return e ? atoi(e) : 0; }();   // tuning aid
    const bool twice = interior_scale == 2.f;
    const float4* table = twice ? plan->lane_adj : plan->lane_fwd;   // (synthesis, doubled) / (analysis)
    B2S_REQUIRE(win == (twice ? plan->swin : plan->awin), "internal: window / table mismatch");
    B2S_REQUIRE(rows * frames < ((int64_t)1 << 30) && samples < ((int64_t)1 << 30) && pad_left < ((int64_t)1 << 30),
                "signal too large for the STFT warp kernel (%lld frames of %lld samples)", (long long)(rows * frames),
                (long long)samples);
    const PipeArgs pa{x, rows, samples, row_stride, pad_left, frames, plan->shift, table, out, ablate, layout,
                      plan->device, stream, feat ? *feat : FeatureArgs{}, twice ? win : plan->awin_half, plan->pair_tw};
    // B2S_FWD_PAIR=0: the 8 x 8 x 8 transform per frame instead of the pair transform per unit of two frames
    const char* pe = getenv("B2S_FWD_PAIR");   // read per launch: one process can compare the two
    const bool pair = !(pe && atoi(pe) == 0);
    if (layout == B2S_SPEC_FEATURE) {
      if (pair) {
        if (plan->shift == 256) return launch_pipe<B2S_SPEC_FEATURE, false, true, 2, 2, false, 2, true>(pa);
        return launch_pipe<B2S_SPEC_FEATURE, false, false, 2, 2, false, 2, true>(pa);
      }
      if (plan->shift == 256) return launch_pipe<B2S_SPEC_FEATURE, false, true>(pa);
      return launch_pipe<B2S_SPEC_FEATURE, false, false>(pa);
    }
    // tuning alternatives of the headline configuration (|Y| epilogue, shift 256): tools/hot_bench.py
    const char* ve = getenv("B2S_FWD_VARIANT");   // read per launch: one process can sweep the shapes
    const int variant = ve ? atoi(ve) : 0;
    if (layout == B2S_SPEC_ABS && plan->shift == 256 && variant != 0 && !pair) {
      switch (variant) {
        case 1: return launch_pipe<B2S_SPEC_ABS, false, true, 2, 3, false, 1>(pa);
        case 2: return launch_pipe<B2S_SPEC_ABS, false, true, 1, 3, false, 2>(pa);
        case 3: return launch_pipe<B2S_SPEC_ABS, false, true, 1, 4, true, 1>(pa);
        default: break;
      }
    }
#define B2S_PIPE_S(L, D) do {                                                                        \
      if (pair) { if (plan->shift == 256) return launch_pipe<L, D, true, 2, 2, false, 2, true>(pa);      \
                  else return launch_pipe<L, D, false, 2, 2, false, 2, true>(pa); }                        \
      if (plan->shift == 256) return launch_pipe<L, D, true>(pa); else return launch_pipe<L, D, false>(pa); } while (0)
    switch (layout) {
      case B2S_SPEC_INTERLEAVED: if (twice) B2S_PIPE_S(B2S_SPEC_INTERLEAVED, true); else B2S_PIPE_S(B2S_SPEC_INTERLEAVED, false); break;
      case B2S_SPEC_CONCAT: if (twice) B2S_PIPE_S(B2S_SPEC_CONCAT, true); else B2S_PIPE_S(B2S_SPEC_CONCAT, false); break;
      case B2S_SPEC_ABS: B2S_PIPE_S(B2S_SPEC_ABS, false); break;
      default: B2S_PIPE_S(B2S_SPEC_LOG1P_ABS, false); break;
    }
#undef B2S_PIPE_S
  } else if (plan->fast) {
    const int64_t want = ceil_div(total, kFwdWarps);
    const int grid = (int)std::min<int64_t>(want, (int64_t)kNumSMs * 3 * 4);
    const bool vec = aligned8(x) && row_stride % 2 == 0 && plan->shift % 2 == 0 && pad_left % 2 == 0;
#define B2S_FWD(L, V)                                                                                  \
    stft1024_forward_kernel<L, V><<<grid, 32 * kFwdWarps, 0, stream>>>(x, rows, samples, row_stride,   \
        pad_left, frames, plan->shift, plan->wlen, win, plan->tw, interior_scale, out)
    switch (layout * 2 + (vec ? 1 : 0)) {
      case 0: B2S_FWD(B2S_SPEC_INTERLEAVED, false); break;
      case 1: B2S_FWD(B2S_SPEC_INTERLEAVED, true); break;
      case 2: B2S_FWD(B2S_SPEC_CONCAT, false); break;
      case 3: B2S_FWD(B2S_SPEC_CONCAT, true); break;
      case 4: B2S_FWD(B2S_SPEC_ABS, false); break;
      case 5: B2S_FWD(B2S_SPEC_ABS, true); break;
      case 6: B2S_FWD(B2S_SPEC_LOG1P_ABS, false); break;
      default: B2S_FWD(B2S_SPEC_LOG1P_ABS, true); break;
    }
#undef B2S_FWD
    B2S_LAUNCH_CHECK("stft1024_forward_kernel");
  } else {
    const int grid = (int)std::min<int64_t>(total, (int64_t)kNumSMs * 64);
    const size_t smem = sizeof(float) * plan->wlen;
    dft_forward_kernel<<<grid, 256, smem, stream>>>(x, rows, samples, row_stride, pad_left, frames,
        plan->size, plan->shift, plan->wlen, win, plan->tw, layout, interior_scale, out);
    B2S_LAUNCH_CHECK("dft_forward_kernel");
  }
  return B2S_OK;
}

bool fused_inverse_ok(const b2s_stft_plan* plan) {
  const int overlap = (plan->wlen + plan->shift - 1) / plan->shift;
  return plan->fast && overlap <= 8;   // hops = slots - overlap + 1 >= 9
}

// the ring kernel: shift 256 (four hops per zero-extended frame) and 16-byte stores of whole hop quarters
bool ring_inverse_ok(const b2s_stft_plan* plan, int64_t crop_left, int64_t samples_out, const float* out) {
  return plan->fast && plan->shift == kRingHop && (crop_left & 3) == 0 && (samples_out & 3) == 0 &&
         (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}
int64_t ring_workspace_bytes(int64_t rows, int64_t frames) {
  int64_t most = 0;
  for (int ns = 1; ns <= 2; ++ns) most = std::max(most, ring_geometry(rows, frames, ns).boundaries);
  return kTicketBytes + most * (int64_t)sizeof(float) * kBoundaryFloats;
}

// inverse-type launch shared by b2s_istft_forward and b2s_stft_backward
int launch_inverse(const b2s_stft_plan* plan, const float* spec, int64_t rows, int64_t frames,
                   int layout, int64_t crop_left, int64_t samples_out, const float* win,
                   float interior_scale /*2: iSTFT, 1: adjoint of STFT*/, float* out, float* scratch,
                   cudaStream_t stream) {
  if (rows * samples_out == 0) return B2S_OK;
  if (frames == 0) {
    B2S_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * rows * samples_out, stream));
    return B2S_OK;
  }
  // B2S_INV_RING: 0 = chunked kernel with halo frames, 1 / 2 = ring kernel with one / two frames per step (default 2)
  const int ring_ns = [] { const char* e = getenv("B2S_INV_RING"); const int v = e ? atoi(e) : 2; return v < 0 || v > 2 ? 2 : v; }();   // per call: tests switch it
  if (fused_inverse_ok(plan) && ring_ns > 0 && ring_inverse_ok(plan, crop_left, samples_out, out) && scratch) {
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(spec) & 7) == 0, "spectrum pointer must be 8-byte aligned");
    const RingGeometry g = ring_geometry(rows, frames, ring_ns);
    B2S_REQUIRE(g.boundaries <= kMaxTickets, "too many rows for the inverse transform's workspace");
    const size_t smem = sizeof(float) * kRingWarps * ring_warp_floats(ring_ns);
    const int variant = (layout == B2S_SPEC_INTERLEAVED ? 0 : 1) + (ring_ns == 2 ? 2 : 0);
    auto kernel = ring_ns == 2
        ? (layout == B2S_SPEC_INTERLEAVED ? istft_ring_kernel<B2S_SPEC_INTERLEAVED, 2> : istft_ring_kernel<B2S_SPEC_CONCAT, 2>)
        : (layout == B2S_SPEC_INTERLEAVED ? istft_ring_kernel<B2S_SPEC_INTERLEAVED, 1> : istft_ring_kernel<B2S_SPEC_CONCAT, 1>);
    static bool configured[4][64] = {};
    if (!configured[variant][plan->device & 63]) {
      B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[variant][plan->device & 63] = true;
    }
    const bool adjoint = interior_scale == 1.f;
    B2S_REQUIRE(win == (adjoint ? plan->awin : plan->swin), "internal: window / table mismatch");
    int* tickets = reinterpret_cast<int*>(scratch);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + kTicketBytes);
    // interleaved rows, two frames per step: the pair transform run backwards (B2S_INV_PAIR=0: rf::irfft_streams<2>)
    const char* ipe = getenv("B2S_INV_PAIR");
    if (ring_ns == 2 && layout == B2S_SPEC_INTERLEAVED && plan->wlen == fft::kSize && !(ipe && atoi(ipe) == 0)) {
      static bool pair_configured[64] = {};
      if (!pair_configured[plan->device & 63]) {
        B2S_CUDA(cudaFuncSetAttribute(istft_pair_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pair_configured[plan->device & 63] = true;
      }
      istft_pair_ring_kernel<<<g.grid, 32 * kRingWarps, smem, stream>>>(spec, rows, frames, g.chunks_per_row, crop_left,
          samples_out, win, plan->pair_tw, 0.5f * interior_scale, out, tickets, partials);
      B2S_LAUNCH_CHECK("istft_pair_ring_kernel");
      return B2S_OK;
    }
    kernel<<<g.grid, 32 * kRingWarps, smem, stream>>>(spec, rows, frames, g.chunks_per_row, crop_left, samples_out,
        adjoint ? plan->lane_inv_ana : plan->lane_inv_syn, 0.5f * interior_scale, out, tickets, partials);
    B2S_LAUNCH_CHECK("istft_ring_kernel");
  } else if (fused_inverse_ok(plan)) {
    B2S_REQUIRE((reinterpret_cast<uintptr_t>(spec) & 7) == 0, "spectrum pointer must be 8-byte aligned");
    static const int ns = [] { const char* e = getenv("B2S_INV_NS"); return e && atoi(e) == 2 ? 2 : 1; }();
    const int overlap = (plan->wlen + plan->shift - 1) / plan->shift;
    const int hops = inv_slots(ns) - overlap + 1;
    // padded samples that can receive output: [crop_left, crop_left + samples_out)
    const int64_t total_hops = ceil_div(crop_left + samples_out, plan->shift);
    const int64_t chunks = ceil_div(total_hops, hops);
    const size_t smem = sizeof(float2) * rf::kTile1 * inv_slots(ns);
    static bool configured[4][64] = {};
    const int variant = (layout == B2S_SPEC_INTERLEAVED ? 0 : 1) + (ns == 2 ? 2 : 0);
    auto kernel = ns == 2
        ? (layout == B2S_SPEC_INTERLEAVED ? istft1024_kernel<B2S_SPEC_INTERLEAVED, 2> : istft1024_kernel<B2S_SPEC_CONCAT, 2>)
        : (layout == B2S_SPEC_INTERLEAVED ? istft1024_kernel<B2S_SPEC_INTERLEAVED, 1> : istft1024_kernel<B2S_SPEC_CONCAT, 1>);
    if (!configured[variant][plan->device & 63]) {
      B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[variant][plan->device & 63] = true;
    }
    const int grid = (int)std::min<int64_t>(rows * chunks, (int64_t)kNumSMs * (ns == 1 ? 3 : 2) * 8);
    // the inverse transform yields S = edge + 2 * interior; the adjoint of the STFT wants edge + 1 * interior
    const bool adjoint = interior_scale == 1.f;
    B2S_REQUIRE(win == (adjoint ? plan->awin : plan->swin), "internal: window / table mismatch");
    kernel<<<grid, 32 * kInvWarps, smem, stream>>>(spec, rows, frames, plan->shift, plan->wlen, overlap, hops,
        chunks, crop_left, samples_out, adjoint ? plan->lane_inv_ana : plan->lane_inv_syn,
        0.5f * interior_scale, out);
    B2S_LAUNCH_CHECK("istft1024_kernel");
  } else {
    B2S_REQUIRE(scratch != nullptr, "inverse transform of this plan needs scratch (b2s_stft_scratch_bytes)");
    const int64_t total = rows * frames;
    const int grid = (int)std::min<int64_t>(total, (int64_t)kNumSMs * 64);
    const size_t smem = sizeof(float2) * plan->bins;
    dft_inverse_frames_kernel<<<grid, 256, smem, stream>>>(spec, total, layout, plan->size, plan->wlen,
        win, plan->tw, interior_scale, scratch);
    B2S_LAUNCH_CHECK("dft_inverse_frames_kernel");
    const int grid2 = (int)std::min<int64_t>(ceil_div(rows * samples_out, 256), (int64_t)kNumSMs * 32);
    overlap_add_gather_kernel<<<grid2, 256, 0, stream>>>(scratch, rows, frames, plan->shift, plan->wlen,
        crop_left, samples_out, out);
    B2S_LAUNCH_CHECK("overlap_add_gather_kernel");
  }
  return B2S_OK;
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

int64_t b2s_stft_frames(int64_t samples, int window_length, int shift, int pad, int fading) {
  const int64_t total = samples + fading_extra(window_length, shift, fading);
  const int64_t num = total - window_length + shift;
  return pad ? -floor_div(-num, shift) : floor_div(num, shift);
}

int64_t b2s_stft_samples(int64_t frames, int window_length, int shift, int fading) {
  return frames * shift + window_length - shift - fading_extra(window_length, shift, fading);
}

int64_t b2s_stft_frame_index(int64_t sample_index, int window_length, int shift, int fading) {
  // PARITY UNPINNED: no reference test pins it; window-centre convention.
  int64_t offset = 0;
  if (fading == 2) offset = (window_length - shift) / 2;
  else if (fading == 1) offset = window_length - shift;
  const int64_t v = floor_div(sample_index + offset - window_length / 2, shift);
  return v < 0 ? 0 : v;
}

int b2s_stft_plan_create(b2s_stft_plan** out, int device, int size, int shift, int window_length,
                         const double* analysis_window, const double* synthesis_window) {
  B2S_REQUIRE(out != nullptr, "plan output pointer is NULL");
  B2S_REQUIRE(size >= 2 && size % 2 == 0, "only even FFT sizes are supported (got %d)", size);
  B2S_REQUIRE(size <= 16384, "FFT size %d exceeds the supported maximum 16384", size);
  B2S_REQUIRE(shift >= 1, "shift must be positive (got %d)", shift);
  B2S_REQUIRE(window_length >= 1 && window_length <= size,
              "window_length must be in [1, size] (got %d, size %d)", window_length, size);
  B2S_REQUIRE(analysis_window && synthesis_window, "window pointers must not be NULL");
  B2S_ON_DEVICE(device);
  std::vector<float> aw(size, 0.f), sw(size, 0.f);
  for (int k = 0; k < window_length; ++k) {
    aw[k] = (float)analysis_window[k];
    sw[k] = (float)synthesis_window[k];
  }
  std::vector<float2> tw(size);
  for (int q = 0; q < size; ++q) {
    const double ang = -2.0 * M_PI * (double)q / (double)size;
    tw[q] = make_float2((float)cos(ang), (float)sin(ang));
  }
  b2s_stft_plan* plan = new b2s_stft_plan();
  plan->device = device; plan->size = size; plan->shift = shift; plan->wlen = window_length;
  plan->bins = size / 2 + 1;
  plan->fast = size == fft::kSize;
  plan->awin = nullptr; plan->swin = nullptr; plan->tw = nullptr;
  plan->lane_fwd = nullptr; plan->lane_adj = nullptr; plan->lane_inv_syn = nullptr; plan->lane_inv_ana = nullptr;
  plan->pair_tw = nullptr; plan->awin_half = nullptr;
  cudaError_t e = cudaMalloc(&plan->awin, sizeof(float) * size);
  if (e == cudaSuccess) e = cudaMalloc(&plan->swin, sizeof(float) * size);
  if (e == cudaSuccess) e = cudaMalloc(&plan->tw, sizeof(float2) * size);
  if (e == cudaSuccess) e = cudaMemcpy(plan->awin, aw.data(), sizeof(float) * size, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(plan->swin, sw.data(), sizeof(float) * size, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(plan->tw, tw.data(), sizeof(float2) * size, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && plan->fast) {
    const size_t bytes = sizeof(float4) * rf::kConstFloat4 * 32;
    std::vector<float4> fwd(rf::kConstFloat4 * 32), adj(rf::kConstFloat4 * 32);
    for (int lane = 0; lane < 32; ++lane) {
      rf::LaneConsts k;
      k.init(tw.data(), aw.data(), lane, 0.5f);
      k.pack(fwd.data());
      k.init(tw.data(), sw.data(), lane, 1.f);
      k.pack(adj.data());
    }
    e = cudaMalloc(&plan->lane_fwd, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&plan->lane_adj, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(plan->lane_fwd, fwd.data(), bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(plan->lane_adj, adj.data(), bytes, cudaMemcpyHostToDevice);
    const size_t ibytes = sizeof(float4) * rf::kInvConstFloat4 * 32;
    std::vector<float4> isyn(rf::kInvConstFloat4 * 32), iana(rf::kInvConstFloat4 * 32);
    for (int lane = 0; lane < 32; ++lane) {
      rf::InvLaneConsts ik;
      ik.init(tw.data(), sw.data(), lane, 1.f);
      ik.pack(isyn.data());
      ik.init(tw.data(), aw.data(), lane, 0.5f);
      ik.pack(iana.data());
    }
    if (e == cudaSuccess) e = cudaMalloc(&plan->lane_inv_syn, ibytes);
    if (e == cudaSuccess) e = cudaMalloc(&plan->lane_inv_ana, ibytes);
    if (e == cudaSuccess) e = cudaMemcpy(plan->lane_inv_syn, isyn.data(), ibytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(plan->lane_inv_ana, iana.data(), ibytes, cudaMemcpyHostToDevice);
    std::vector<float2> ptw(32 * 32);
    for (int q = 0; q < 32; ++q)
      for (int lane = 0; lane < 32; ++lane) ptw[q * 32 + lane] = tw[(lane * q) & 1023];
    std::vector<float> half(size);
    for (int i = 0; i < size; ++i) half[i] = 0.5f * aw[i];
    if (e == cudaSuccess) e = cudaMalloc(&plan->awin_half, sizeof(float) * size);
    if (e == cudaSuccess) e = cudaMemcpy(plan->awin_half, half.data(), sizeof(float) * size, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&plan->pair_tw, sizeof(float2) * 32 * 32);
    if (e == cudaSuccess) e = cudaMemcpy(plan->pair_tw, ptw.data(), sizeof(float2) * 32 * 32, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    set_error("stft plan allocation failed: %s", cudaGetErrorString(e));
    b2s_stft_plan_destroy(plan);
    return B2S_ERR_CUDA;
  }
  *out = plan;
  return B2S_OK;
}

int b2s_stft_plan_destroy(b2s_stft_plan* plan) {
  if (!plan) return B2S_OK;
  DeviceGuard guard(plan->device);
  cudaFree(plan->awin); cudaFree(plan->swin); cudaFree(plan->tw);
  cudaFree(plan->lane_fwd); cudaFree(plan->lane_adj);
  cudaFree(plan->lane_inv_syn); cudaFree(plan->lane_inv_ana);
  cudaFree(plan->pair_tw); cudaFree(plan->awin_half);
  delete plan;
  return B2S_OK;
}

int b2s_stft_plan_is_fast(const b2s_stft_plan* plan) { return plan && plan->fast; }

int64_t b2s_stft_scratch_bytes(const b2s_stft_plan* plan, int64_t rows, int64_t frames) {
  if (!plan) return 0;
  if (fused_inverse_ok(plan)) return plan->shift == kRingHop ? ring_workspace_bytes(rows, frames) : 0;
  return (int64_t)sizeof(float) * rows * frames * plan->wlen;
}

int b2s_stft_forward(const b2s_stft_plan* plan, const float* signal, int64_t rows, int64_t samples,
                     int64_t row_stride, int64_t pad_left, int64_t frames, int layout, float* spec,
                     b2s_stream stream) {
  if (int rc = check_plan(plan)) return rc;
  B2S_ON_DEVICE(plan->device);
  B2S_REQUIRE(layout >= 0 && layout <= 3, "unknown spectrum layout %d", layout);
  B2S_REQUIRE(rows >= 0 && samples >= 0 && frames >= 0 && pad_left >= 0, "negative extent");
  B2S_REQUIRE(rows * frames == 0 || (signal && spec), "NULL device pointer");
  return launch_forward(plan, signal, rows, samples, row_stride, pad_left, frames, layout, plan->awin,
                        1.f, spec, (cudaStream_t)stream);
}

int b2s_istft_backward(const b2s_stft_plan* plan, const float* grad_signal, int64_t rows,
                       int64_t samples_out, int64_t crop_left, int64_t frames, int layout,
                       float* grad_spec, b2s_stream stream) {
  if (int rc = check_plan(plan)) return rc;
  B2S_ON_DEVICE(plan->device);
  B2S_REQUIRE(layout == B2S_SPEC_INTERLEAVED || layout == B2S_SPEC_CONCAT, "layout must be complex");
  B2S_REQUIRE(rows * frames == 0 || (grad_signal && grad_spec), "NULL device pointer");
  // d signal / d Re,Im = c_f * STFT with the synthesis window (c_f = 2 for interior bins)
  return launch_forward(plan, grad_signal, rows, samples_out, samples_out, crop_left, frames, layout,
                        plan->swin, 2.f, grad_spec, (cudaStream_t)stream);
}

int b2s_istft_forward(const b2s_stft_plan* plan, const float* spec, int64_t rows, int64_t frames,
                      int layout, int64_t crop_left, int64_t samples_out, float* signal,
                      float* scratch, b2s_stream stream) {
  if (int rc = check_plan(plan)) return rc;
  B2S_ON_DEVICE(plan->device);
  B2S_REQUIRE(layout == B2S_SPEC_INTERLEAVED || layout == B2S_SPEC_CONCAT, "layout must be complex");
  B2S_REQUIRE(rows * samples_out == 0 || (spec || frames == 0) && signal, "NULL device pointer");
  return launch_inverse(plan, spec, rows, frames, layout, crop_left, samples_out, plan->swin, 2.f,
                        signal, scratch, (cudaStream_t)stream);
}

int b2s_stft_backward(const b2s_stft_plan* plan, const float* grad_spec, int64_t rows, int64_t frames,
                      int layout, int64_t pad_left, int64_t samples, float* grad_signal,
                      float* scratch, b2s_stream stream) {
  if (int rc = check_plan(plan)) return rc;
  B2S_ON_DEVICE(plan->device);
  B2S_REQUIRE(layout == B2S_SPEC_INTERLEAVED || layout == B2S_SPEC_CONCAT, "layout must be complex");
  B2S_REQUIRE(rows * samples == 0 || (grad_spec || frames == 0) && grad_signal, "NULL device pointer");
  return launch_inverse(plan, grad_spec, rows, frames, layout, pad_left, samples, plan->awin, 1.f,
                        grad_signal, scratch, (cudaStream_t)stream);
}

// ---- feature epilogues ------------------------------------------------------------
struct b2s_mel {
  int device, bins, filters;
  int *lo, *len, *woff;
  float* weights;
};

int b2s_mel_create(b2s_mel** out, int device, int bins, int filters, const float* basis) {
  B2S_REQUIRE(out && basis, "NULL pointer");
  B2S_REQUIRE(bins >= 1 && filters >= 1 && filters <= 4096, "bad filterbank extents (%d bins, %d filters)", bins, filters);
  B2S_ON_DEVICE(device);
  // support of every filter (column of basis [bins][filters]): first .. last non-zero bin
  std::vector<int> lo(filters, 0), len(filters, 0), woff(filters, 0);
  std::vector<float> w;
  for (int m = 0; m < filters; ++m) {
    int first = -1, last = -1;
    for (int f = 0; f < bins; ++f)
      if (basis[(size_t)f * filters + m] != 0.f) { if (first < 0) first = f; last = f; }
    woff[m] = (int)w.size();
    if (first >= 0) {
      lo[m] = first; len[m] = last - first + 1;
      for (int f = first; f <= last; ++f) w.push_back(basis[(size_t)f * filters + m]);
    }
  }
  if (w.empty()) w.push_back(0.f);
  b2s_mel* mel = new b2s_mel();
  mel->device = device; mel->bins = bins; mel->filters = filters;
  B2S_CUDA(cudaMalloc(&mel->lo, sizeof(int) * filters));
  B2S_CUDA(cudaMalloc(&mel->len, sizeof(int) * filters));
  B2S_CUDA(cudaMalloc(&mel->woff, sizeof(int) * filters));
  B2S_CUDA(cudaMalloc(&mel->weights, sizeof(float) * w.size()));
  B2S_CUDA(cudaMemcpy(mel->lo, lo.data(), sizeof(int) * filters, cudaMemcpyHostToDevice));
  B2S_CUDA(cudaMemcpy(mel->len, len.data(), sizeof(int) * filters, cudaMemcpyHostToDevice));
  B2S_CUDA(cudaMemcpy(mel->woff, woff.data(), sizeof(int) * filters, cudaMemcpyHostToDevice));
  B2S_CUDA(cudaMemcpy(mel->weights, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  *out = mel;
  return B2S_OK;
}

int b2s_mel_destroy(b2s_mel* mel) {
  if (!mel) return B2S_OK;
  DeviceGuard guard(mel->device);
  cudaFree(mel->lo); cudaFree(mel->len); cudaFree(mel->woff); cudaFree(mel->weights);
  delete mel;
  return B2S_OK;
}

int b2s_stft_features(const b2s_stft_plan* plan, const float* signal, int64_t rows, int64_t samples,
                      int64_t row_stride, int64_t pad_left, int64_t frames, float power, float scale,
                      int log_kind, float eps, const b2s_mel* mel, float* features, b2s_stream stream) {
  if (int rc = check_plan(plan)) return rc;
  B2S_REQUIRE(plan->fast && plan->wlen == fft::kSize && plan->shift % 4 == 0 && plan->shift <= fft::kSize,
              "the fused feature epilogue exists for size 1024 / window_length 1024 / shift %% 4 == 0 plans only "
              "(got size %d, window_length %d, shift %d)", plan->size, plan->wlen, plan->shift);
  B2S_REQUIRE(power > 0.f, "power must be positive (got %g)", (double)power);
  B2S_REQUIRE(log_kind >= B2S_LOG_NONE && log_kind <= B2S_LOG_2, "unknown logarithm %d", log_kind);
  B2S_REQUIRE(!mel || (mel->bins == plan->bins && mel->device == plan->device),
              "filterbank was built for %d bins on device %d", mel ? mel->bins : 0, mel ? mel->device : 0);
  B2S_REQUIRE(rows >= 0 && samples >= 0 && frames >= 0 && pad_left >= 0, "negative extent");
  B2S_REQUIRE(rows * frames == 0 || (signal && features), "NULL device pointer");
  B2S_ON_DEVICE(plan->device);
  FeatureArgs feat{0.5f * power, scale, log_kind, eps, mel ? mel->filters : 0, mel ? mel->lo : nullptr,
                   mel ? mel->len : nullptr, mel ? mel->woff : nullptr, mel ? mel->weights : nullptr};
  return launch_forward(plan, signal, rows, samples, row_stride, pad_left, frames, B2S_SPEC_FEATURE, plan->awin, 1.f,
                        features, (cudaStream_t)stream, &feat);
}

}  // extern "C"
