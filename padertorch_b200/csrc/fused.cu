// The fused north-star kernel: STFT -> |.| -> mask (*) |Y| -> K x K SSE -> permutation search, with the
// target spectra |STFT(s_k)| recomputed in registers instead of being written to and re-read from HBM.
// Equals  pit_loss(mask * Y_abs[:, None, :], X_abs, axis=-2)  with  X_abs = |STFT(s)|
// (padertorch/contrib/examples/source_separation/pit/data.py:49-77, pit/model.py:117-128,
//  padertorch/ops/losses/source_separation.py:34-124).
// Algorithmic HBM bytes per utterance: 4T(1+K) + 4MFK; reading the already
// materialised |Y| instead of recomputing it trades 4T for 4MF bytes against one FFT per frame.
//
// Work decomposition: the frame positions of the whole batch form one list; every warp of a persistent grid
// is an independent pipeline that owns a contiguous range of it (equal work per warp, at most a few example
// boundaries per warp).  Per position the warp's elected lane starts TMA bulk copies (cp.async.bulk ->
// mbarrier): the 1024-sample frames of the K sources (re-started for the next position as soon as pass 1
// holds the current samples in registers) and the position's mask rows [K][513] and |Y| row (re-started
// after the SSE).  The K transforms run two at a time (rfft_packed.cuh); their magnitudes never leave the
// registers.  No block-wide barrier and no per-thread staging code exist in the steady state.
// At an example boundary the warp flushes its K x K partial sums; the last warp of an example (ticket)
// folds the partials in slot order and searches the K! permutations (itertools order, first minimum wins).
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "fft1024.cuh"
#include "rfft_packed.cuh"
#include "cfft_pair.cuh"
#include "tma.cuh"
#include "stft_plan.cuh"
#include "perm.cuh"

using namespace b2s;
using namespace b2s::tma;

namespace {

// Resident pipelines per SM.  A pipeline (= warp) needs shared memory for its NT frames, NS exchange tiles and the
// mask / |Y| rows of a position, and registers for NS interleaved transforms.  Default: NS = 2 (two transforms
// interleaved, <= 255 registers), 8 warps per SM.  Measured alternatives at the north-star shape (B200, round 2,
// profiles/r2_fused_shapes.txt): one transform at a time at <= 168 registers and 12 warps per SM (19 040 bytes per
// warp at K = 2) in one, two or three CTAs takes 58.3 / 60.0 / 59.6 us against 50.5 us, 10 warps at <= 200
// registers 55.7 us -- the SM's time per position is the same 0.36 us whatever the number of resident warps (the
// transform alone: 0.78 ns per frame at 8 warps per SM, 0.73 at 12, tools/ubench/rfft_rate.cu): the kernel is bound
// by the work per position on the shared-memory and FP32 pipes, not by latency hiding.
struct FusedShape { int warps, ctas, ns; };
__host__ __device__ constexpr FusedShape fused_shape(int K, bool recompute, int variant) {
  // variant 0 = default of this (K, recompute); others are tuning alternatives (B2S_FUSED_VARIANT, K <= 2 only)
  if (K <= 2 && !recompute) {
    switch (variant) {
      case 1: return FusedShape{12, 1, 1};
      case 2: return FusedShape{6, 2, 1};
      case 3: return FusedShape{4, 3, 1};
      case 4: return FusedShape{10, 1, 1};   // <= 200 registers
      default: return FusedShape{4, 2, 2};
    }
  }
  if (K <= 2) return FusedShape{4, 2, 2};
  if (K == 3) return FusedShape{recompute ? 6 : 7, 1, 2};
  return FusedShape{3, 2, 2};
}
constexpr int kFusedVariants = 5;

struct FusedGrid {
  int grid;            // persistent CTAs
  int64_t warps;       // pipelines = grid * warps per CTA
  int64_t total;       // batch * frames positions (dense enumeration over `frames`)
  int slots;           // partial-sum slots per example
};

FusedGrid fused_grid(int64_t batch, int64_t frames, FusedShape shape) {
  FusedGrid g;
  g.total = batch * std::max<int64_t>(1, frames);
  // B2S_FUSED_SMS (tuning aid): use fewer SMs -- separates per-SM limits from chip-wide contention
  static const int sms = [] { const char* e = getenv("B2S_FUSED_SMS"); const int v = e ? atoi(e) : 0; return v > 0 && v < kNumSMs ? v : kNumSMs; }();
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(g.total, shape.warps),
                                                       (int64_t)sms * shape.ctas));
  g.warps = std::min<int64_t>((int64_t)g.grid * shape.warps, g.total);   // surplus warps of the last CTA idle
  g.slots = (int)(ceil_div(std::max<int64_t>(1, frames) * g.warps, g.total) + 2);
  return g;
}
// the workspace must hold the partial sums of whichever shape is launched
int fused_max_slots(int64_t batch, int64_t frames, int K) {
  int slots = 0;
  for (int r = 0; r < 2; ++r)
    for (int v = 0; v < kFusedVariants; ++v) slots = std::max(slots, fused_grid(batch, frames, fused_shape(K, r != 0, v)).slots);
  return slots;
}

// first position of warp w: floor(w * total / warps); warp owning position x: ceil((x + 1) * warps / total) - 1
__device__ __forceinline__ int64_t range_start(int64_t w, int64_t total, int64_t warps) {
  return w * total / warps;
}
__device__ __forceinline__ int64_t owner_of(int64_t x, int64_t total, int64_t warps) {
  return ((x + 1) * warps + total - 1) / total - 1;
}

// floats of a warp's mask / |Y| landing area: [mask rows K * F + 8 | |Y| row F + 8], both 16-byte aligned
__host__ __device__ constexpr int mask_area_floats(int K) { return ((K * 513 + 8 + 3) / 4) * 4; }
__host__ __device__ constexpr int row_area_floats(int K) { return mask_area_floats(K) + ((513 + 8 + 3) / 4) * 4; }

// magnitudes of one slot pair (A side, B side)
__device__ __forceinline__ float2 mag2(float2 ya, float2 yb) {
  return make_float2(fft::sqrt_approx(fmaf(ya.x, ya.x, ya.y * ya.y)),
                     fft::sqrt_approx(fmaf(yb.x, yb.x, yb.y * yb.y)));
}

// All K! (K <= 4: at most 24) assignments evaluated by the lanes of ONE warp; lane 0 receives the winner.
__device__ __forceinline__ void warp_search_permutations(const double* cost, int K, int lane, double& best_value,
                                                         int* best_perm) {
  const int total = factorial(K);
  double val = 0.0;
  int idx = 0x7fffffff;
  if (lane < total) {
    int p[B2S_MAX_SOURCES];
    unrank_permutation(lane, K, p);
    for (int k = 0; k < K; ++k) val += cost[p[k] * K + k];
    idx = lane;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, val, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (oi != 0x7fffffff && (idx == 0x7fffffff || candidate_better(ov, oi, val, idx))) { val = ov; idx = oi; }
  }
  best_value = val;
  if (lane == 0) unrank_permutation(idx, K, best_perm);
}

// One frame position = (K [+1]) transforms whose magnitudes stay in registers as packed (A side, B side)
// pairs, then the K x K SSE of 9 packed bin pairs per lane (slot 8 = DC / Nyquist, live in lane 0 only).
// RING (shift 256 only): the frames of a position are not copied whole.  Every transform keeps a ring of five hops
// of 256 samples; a warp walks consecutive positions of an utterance, so position m + 1 needs ONE new hop per
// signal (1 KB instead of 4 KB: a quarter of the L2 -> shared-memory traffic and no re-fetch of overlapping frames
// from DRAM), and that hop's slot is not part of the frame being transformed: it is requested a whole position
// ahead.  Only the first position of a range / an utterance fills four hops.
constexpr int kHop = 256, kRingHops = 5;
// Tuning aids (per-warp trace stamps, ablation bits) are compiled in only with -DB2S_TUNING=1 (B2S_TUNING=1 python -m
// padertorch_b200.build --force): in the production kernel they cost ~3 % (branches inside the position loop).
#ifndef B2S_TUNING
#define B2S_TUNING 0
#endif
constexpr bool kTuning = B2S_TUNING != 0;
template <int K, bool RECOMPUTE_Y, int WARPS, int CTAS, int NS, bool RING>
__global__ void __launch_bounds__(32 * WARPS, CTAS)
stft_pit_fused_kernel(const float* __restrict__ mixture, const float* __restrict__ yabs,
                      const float* __restrict__ sources, const float* __restrict__ mask,
                      const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames,
                      int shift, int64_t pad_left, const float4* __restrict__ lane_table, int slots,
                      double* __restrict__ partial, int* __restrict__ counters, float* __restrict__ loss,
                      int32_t* __restrict__ perm, double* __restrict__ sse, unsigned long long* __restrict__ trace,
                      int ablate_arg /* tuning only: 1 no SSE, 2 no square roots, 4 no row copies, 8 no frame copies */) {
  const int ablate = kTuning ? ablate_arg : 0;
  if (!kTuning) trace = nullptr;
  constexpr int NV = K * K;
  constexpr int F = rf::kBins;
  constexpr int kFusedWarps = WARPS;
  constexpr int NT = RECOMPUTE_Y ? K + 1 : K;   // transforms per position; with RECOMPUTE_Y the mixture is first
  constexpr int kSig = RING ? kRingHops * kHop : rf::kSize;   // floats per staged signal
  constexpr int kWarpFloats = NT * kSig + 2 * NS * rf::kTile1 + row_area_floats(K);
  extern __shared__ __align__(16) float smem[];   // per warp: [NT][1024] frames, NS exchange tiles, mask / |Y| rows
  __shared__ __align__(8) uint64_t bars[kFusedWarps][2];
  __shared__ double totals_sm[kFusedWarps][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // B2S_FUSED_TRACE (tuning aid): %globaltimer of five events per warp
  auto stamp = [&](int what) {
    if (trace && lane == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[((size_t)blockIdx.x * kFusedWarps + warp) * 8 + what] = t;
    }
  };
  stamp(0);
  if (trace && lane == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[((size_t)blockIdx.x * kFusedWarps + warp) * 8 + 6] = smid;
  }
  float* sig = smem + warp * kWarpFloats;                          // frame of transform t at sig + t * 1024
  float2* tile = reinterpret_cast<float2*>(sig + NT * kSig);
  float* rows_area = sig + NT * kSig + 2 * NS * rf::kTile1;
  uint64_t* bar_sig = &bars[warp][0];
  uint64_t* bar_rows = &bars[warp][1];
  if (lane == 0) {
    mbar_init(bar_sig, 1);
    mbar_init(bar_rows, 1);
    fence_mbar_init();
  }
  __syncwarp();
  // a kernel launched with programmatic stream serialization behind this one (normally the next STFT front-end,
  // which waits for this kernel before it touches any data) may be scheduled as SMs free up
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  rf::LaneConsts k;
  k.load(lane_table, lane);
  // slot p holds bin (p < 4 ? k0 : k4) + 64 p on the A side and 512 minus that on the B side
  const int k0 = rf::bin_a(lane, 0), k4 = rf::bin_a(lane, 4) - 256;
  const bool first = lane == 0;

  const int64_t total = batch * frames;
  const int64_t nwarps = min((int64_t)gridDim.x * kFusedWarps, total);   // every pipeline owns >= 1 position
  const int64_t gw = (int64_t)blockIdx.x * kFusedWarps + warp;
  if (gw >= nwarps) return;
  const int64_t p_begin = range_start(gw, total, nwarps), p_end = range_start(gw + 1, total, nwarps);

  unsigned sig_phase = 0, rows_phase = 0;
  bool sig_by_tma = false, sig_async = false;
  int off_m = 0, off_y = 0;     // float offset of a row inside its landing area (source misalignment / 4)

  auto frames_of = [&](int64_t b) { return meta ? meta[2 * b + 1] : frames; };
  // row t of a position: t < K source t, t == K the mixture; order of the transforms: (mixture,) sources
  auto signal_row = [&](int64_t b, int t) -> const float* {
    const int r = RECOMPUTE_Y ? (t == 0 ? K : t - 1) : t;
    return r < K ? sources + (b * K + r) * samples : mixture + b * samples;
  };
  // Per-example context (warp uniform), recomputed only when the example changes: the steady-state loop has no
  // 64-bit multiplications or divisions.  Sample offsets fit 32 bits (checked by the launcher).
  int64_t ctx_b = -1;
  int ctx_T = 0, ctx_M = 0;
  bool ctx_a16 = false;
  const float* ctx_row[NT];
  const float* ctx_mask = nullptr;
  const float* ctx_y = nullptr;
  auto set_ctx = [&](int64_t b) {
    if (b == ctx_b) return;
    ctx_b = b;
    ctx_T = (int)(meta ? meta[2 * b] : samples);
    ctx_M = (int)frames_of(b);
    ctx_a16 = true;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      ctx_row[t] = signal_row(b, t);
      ctx_a16 = ctx_a16 && (reinterpret_cast<uintptr_t>(ctx_row[t]) & 15) == 0;
    }
    ctx_mask = mask + b * frames * (K * F);
    ctx_y = RECOMPUTE_Y ? nullptr : yabs + b * frames * F;
  };
  const int pad = (int)pad_left;
  // start the copy of the NT frames of position q = (b, m) (TMA, or zero-filling cp.async for frames that touch
  // the zero padding at the signal's ends or are not 16-byte aligned)
  int64_t ring_b = -1;   // RING: the position whose four hops the rings hold (or will hold once the copies land)
  int ring_m = -2;
  auto start_signals = [&](int64_t q, int64_t b, int m) {
    if (q >= p_end) return;
    set_ctx(b);
    if (m >= ctx_M) { sig_by_tma = false; return; }
    if ((ablate & 8) && q != p_begin) { sig_by_tma = false; return; }
    if (RING) {
      // hops m .. m + 3 of example b must be resident; a continuation needs only hop m + 3
      const bool cont = b == ring_b && m == ring_m + 1;
      const int h0 = cont ? m + 3 : m, h1 = m + 4;
      ring_b = b; ring_m = m;
      sig_async = false;
      {   // the steady state: one hop, inside the signal, 16-byte aligned (pad_left % 4 == 0: launcher)
        const int s0 = h0 * kHop - pad;
        if (cont && ctx_a16 && s0 >= 0 && s0 + kHop <= ctx_T) {
          sig_by_tma = true;
          if (lane == 0) {
            float* dst = sig + ((unsigned)h0 % kRingHops) * kHop;
            mbar_expect_tx(bar_sig, (unsigned)(NT * kHop * 4));
#pragma unroll
            for (int t = 0; t < NT; ++t) bulk_g2s(dst + t * kSig, ctx_row[t] + s0, kHop * 4u, bar_sig);
          }
          return;
        }
      }
      sig_async = true;
      int nbulk = 0;
      for (int h = h0; h < h1; ++h) {
        const int s0 = h * kHop - pad;
        nbulk += (ctx_a16 && (s0 & 3) == 0 && s0 >= 0 && s0 + kHop <= ctx_T) ? 1 : 0;
      }
      sig_by_tma = nbulk > 0;
      if (nbulk > 0 && lane == 0) mbar_expect_tx(bar_sig, (unsigned)(nbulk * NT * kHop * 4));
      for (int h = h0; h < h1; ++h) {
        const int s0 = h * kHop - pad;
        float* dst = sig + (h % kRingHops) * kHop;
        if (ctx_a16 && (s0 & 3) == 0 && s0 >= 0 && s0 + kHop <= ctx_T) {
          if (lane == 0) {
#pragma unroll
            for (int t = 0; t < NT; ++t) bulk_g2s(dst + t * kSig, ctx_row[t] + s0, kHop * 4u, bar_sig);
          }
        } else {   // hop touches the zero padding / the signal's end, or is not 16-byte aligned: zero-filling cp.async
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const float* xr = ctx_row[t];
            for (int i = lane; i < kHop; i += 32) {
              const int n = s0 + i;
              const bool ok = n >= 0 && n < ctx_T;
              fft::cp_async_4_zfill(dst + t * kSig + i, ok ? xr + n : xr, ok ? 4 : 0);
            }
          }
        }
      }
      fft::cp_async_commit();
      return;
    }
    const int s0 = m * shift - pad;
    const bool a16 = ctx_a16 && (s0 & 3) == 0;
    const bool bulk = a16 && s0 >= 0 && s0 + rf::kSize <= ctx_T;
    sig_by_tma = bulk;
    if (bulk) {
      if (lane == 0) {
        // (no fence.proxy.async: the area was only READ through the generic proxy, every lane's loads were consumed
        // before the __syncwarp() that precedes this call, and the copy's writes arrive a memory latency later)
        mbar_expect_tx(bar_sig, NT * rf::kSize * 4u);
#pragma unroll
        for (int t = 0; t < NT; ++t) bulk_g2s(sig + t * kSig, ctx_row[t] + s0, rf::kSize * 4u, bar_sig);
      }
    } else {
      // just as asynchronous as the bulk copy (cp.async.wait_group before pass 1); 16-byte units when possible
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const float* xr = ctx_row[t];
        if (a16) {
          for (int c = lane; c < rf::kSize / 4; c += 32) {
            const int n = s0 + 4 * c;
            const int bytes = n < 0 ? 0 : max(0, min(4, ctx_T - n)) * 4;
            fft::cp_async_16(sig + t * kSig + 4 * c, bytes ? xr + n : xr, bytes);
          }
        } else {
          for (int i = lane; i < rf::kSize; i += 32) {
            const int n = s0 + i;
            const bool ok = n >= 0 && n < ctx_T;
            fft::cp_async_4_zfill(sig + t * kSig + i, ok ? xr + n : xr, ok ? 4 : 0);
          }
        }
      }
      fft::cp_async_commit();
    }
  };
  // Start the copy of the mask rows [K][F] (contiguous) and the |Y| row of position q.  A bulk copy needs
  // 16-byte aligned addresses and sizes while rows start at multiples of 4 bytes: the enclosing aligned range is
  // copied and the row found at the source's misalignment inside the landing area.  The at most 12 bytes before
  // / after a row belong to the neighbouring rows or, at the very ends of the tensor, to the same (>= 256-byte
  // granular) allocation -- an address that is not 16-byte aligned is never at its edge.
  auto start_rows = [&](int64_t q, int64_t b, int m) {
    if (q >= p_end) return;
    set_ctx(b);
    if (m >= ctx_M) return;
    if ((ablate & 4) && q != p_begin) return;
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
        bulk_g2s(rows_area + mask_area_floats(K), reinterpret_cast<const void*>(ay & ~(uintptr_t)15), bytes_y, bar_rows);
    }
  };

  float2 acc[NV];   // (A-side sum, B-side sum)
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float2(0.f, 0.f);

  // flush the warp's partial sums of example b (whole warp)
  auto flush = [&](int64_t b) {
    double mine[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      mine[i] = (double)warp_sum(acc[i].x + acc[i].y);
      acc[i] = make_float2(0.f, 0.f);
    }
    const int64_t first_owner = owner_of(b * frames, total, nwarps);
    const int64_t last_owner = owner_of((b + 1) * frames - 1, total, nwarps);
    const int slot = (int)(gw - first_owner), nparts = (int)(last_owner - first_owner + 1);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) partial[(b * slots + slot) * NV + i] = mine[i];
      __threadfence();
    }
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(counters + b, 1) == nparts - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {   // warp-uniform
      __threadfence();
      double* totals = totals_sm[warp];
      // fold the partials in a fixed order: lane (c0, i) sums slots c0, c0 + per, ... of value i (independent
      // loads in flight instead of one dependent chain), then the `per` lanes of a value are added in lane order
      constexpr int per = 32 / NV;
      const int vi = lane % NV, c0 = lane / NV;
      double s = 0.0;
      if (c0 < per) {
        // four loads in flight per lane (the values come from L2: one round trip instead of one per slot)
        const volatile double* p = partial + b * slots * NV + vi;
        for (int c = c0; c < nparts; c += 4 * per) {
          const double d0 = p[(int64_t)c * NV];
          const double d1 = c + per < nparts ? p[(int64_t)(c + per) * NV] : 0.0;
          const double d2 = c + 2 * per < nparts ? p[(int64_t)(c + 2 * per) * NV] : 0.0;
          const double d3 = c + 3 * per < nparts ? p[(int64_t)(c + 3 * per) * NV] : 0.0;
          s += d0; s += d1; s += d2; s += d3;
        }
      }
#pragma unroll
      for (int j = 1; j < per; ++j) {
        const double o = __shfl_sync(0xffffffffu, s, (vi + j * NV) & 31);
        if (lane < NV) s += o;
      }
      if (lane < NV) {
        totals[lane] = s;
        sse[b * NV + lane] = s;
      }
      __syncwarp();
      double best;
      int bp[B2S_MAX_SOURCES];
      warp_search_permutations(totals, K, lane, best, bp);
      if (lane == 0) {
        loss[b] = (float)(best / ((double)frames_of(b) * (double)K * (double)F));
        for (int kk = 0; kk < K; ++kk) perm[b * K + kk] = bp[kk];
        counters[b] = 0;
      }
      __syncwarp();
    }
  };

  // (example, frame) of the current and of the next position are tracked incrementally: no 64-bit divisions in
  // the loop
  int64_t b = p_begin / frames;
  int m = (int)(p_begin - b * frames);
  const int frames_i = (int)frames;
  // Barriers, the constant table and the index arithmetic above are independent of the preceding kernel; every
  // caller tensor (source / mixture waveforms, masks, |Y|) may have been written by it: nothing of them is
  // requested before this wait.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  start_signals(p_begin, b, m);
  start_rows(p_begin, b, m);
  stamp(1);
  int64_t b_cur = p_begin < p_end ? b : -1;
  for (int64_t q = p_begin; q < p_end; ++q) {
    if (q == p_begin + 1) stamp(2);
    if (q == p_begin + 2) stamp(3);
    // next position
    int64_t bn = b;
    int mn = m + 1;
    if (mn == frames_i) { mn = 0; ++bn; }
    if (b != b_cur) {   // warp-uniform
      flush(b_cur);
      b_cur = b;
    }
    set_ctx(b);
    if (m >= ctx_M) {   // position beyond this example's length (ragged batch): nothing to do
      start_signals(q + 1, bn, mn);
      start_rows(q + 1, bn, mn);
      b = bn; m = mn;
      continue;
    }
    if (sig_by_tma) {
      mbar_wait(bar_sig, sig_phase);
      sig_phase ^= 1;
    }
    if (RING ? sig_async : !sig_by_tma) {   // a ring fill may mix bulk copies and zero-filling cp.async
      fft::cp_async_wait_all();
      __syncwarp();
    }
    // RING: the hop the NEXT position adds lies outside the frame being transformed: request it now, a whole position
    // ahead; a position that starts another utterance (or follows a skipped one) refills from the hook below
    int hop_off[4];
    {
      const int base = (int)((unsigned)m % kRingHops);
#pragma unroll
      for (int j = 0; j < 4; ++j) hop_off[j] = (base + j >= kRingHops ? base + j - kRingHops : base + j) * kHop;
    }
    const bool cont_next = RING && bn == b && mn < ctx_M && q + 1 < p_end;
    if (cont_next) start_signals(q + 1, bn, mn);
    // magnitudes as (A side, B side) pairs; slot 8 = (|DC|, |Nyquist|), zero outside lane 0; lane 0's slot 7
    // holds bin 256 on both sides: its B copy is zeroed (and so is the matching mask load)
    auto magnitudes = [&](const float2 (&ya)[8], const float2 (&yb)[8], float ydc, float ynyq, float2 (&x)[9]) {
#pragma unroll
      for (int p = 0; p < 8; ++p) x[p] = (ablate & 2) ? make_float2(ya[p].x + ya[p].y, yb[p].x + yb[p].y) : mag2(ya[p], yb[p]);
      if (first) x[7].y = 0.f;
      x[8] = first ? make_float2(fabsf(ydc), fabsf(ynyq)) : make_float2(0.f, 0.f);
    };
    float2 x[NT][9];   // with RECOMPUTE_Y x[0] = |Y|, sources follow
#pragma unroll(K <= 2 ? 2 : 1)
    for (int t = 0; NS == 2 && t + 1 < NT; t += 2) {
      float2 ya[2][8], yb[2][8];
      float ydc[2], ynyq[2];
      // the frames of the LAST transforms are in registers after pass 1: start the next position's copy
      auto next_copy = [&]() { if (t + 2 >= NT && !cont_next) start_signals(q + 1, bn, mn); };
      rf::rfft_streams<2, false, false>(sig + t * kSig, kSig, tile, k, ya, yb, ydc, ynyq, 0, next_copy,
                                        RING ? hop_off : nullptr);
      magnitudes(ya[0], yb[0], ydc[0], ynyq[0], x[t]);
      magnitudes(ya[1], yb[1], ydc[1], ynyq[1], x[t + 1]);
    }
    // one transform at a time: all of them (NS == 1) or the odd one out (NS == 2)
#pragma unroll
    for (int t = (NS == 2 ? (NT & ~1) : 0); t < NT; ++t) {
      float2 ya[1][8], yb[1][8];
      float ydc[1], ynyq[1];
      auto next_copy = [&]() { if (t == NT - 1 && !cont_next) start_signals(q + 1, bn, mn); };
      rf::rfft_streams<1, false, false>(sig + t * kSig, 0, tile, k, ya, yb, ydc, ynyq, 0, next_copy,
                                        RING ? hop_off : nullptr);
      magnitudes(ya[0], yb[0], ydc[0], ynyq[0], x[t]);
    }
    constexpr int XS = RECOMPUTE_Y ? 1 : 0;   // x[XS + j] = |STFT(s_j)|

    // ---- SSE of this frame: e_i = mask_i * |Y| against every source magnitude
    // (waiting for the rows earlier -- inside the last transform, before its pass 3, so that the row loads could be
    // scheduled into the pass-3 arithmetic -- was measured slower: 49.4 against 48.1 us, profiles/r2_fused_experiments.txt)
    if (!(ablate & 4) || q == p_begin) {
      mbar_wait(bar_rows, rows_phase);   // the rows of this frame have landed
      rows_phase ^= 1;
    }
    const float* mrow = rows_area + off_m;
    const float* yrow = rows_area + mask_area_floats(K) + off_y;
    if (ablate & 1) {
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int p = 0; p < 9; ++p) acc[t % NV] = rf::add2(acc[t % NV], x[t][p]);
    } else
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
    start_rows(q + 1, bn, mn);
    b = bn; m = mn;
  }
  stamp(4);
  if (b_cur >= 0) flush(b_cur);
  stamp(5);
}

#include "fused_ws.cuh"
#include "fused_pair.cuh"

template <int K, bool RECOMPUTE, int VARIANT>
int launch_fused_shape(const b2s_stft_plan* plan, const float* mixture, const float* yabs, const float* sources,
                       const float* mask, const int64_t* meta, int64_t batch, int64_t samples, int64_t frames,
                       int64_t pad_left, float* loss, int32_t* perm, double* sse, void* workspace,
                       cudaStream_t stream) {
  constexpr FusedShape shape = fused_shape(K, RECOMPUTE, VARIANT);
  const FusedGrid g = fused_grid(batch, frames, shape);
  constexpr int kFusedWarps = shape.warps;
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  // the transform reads its frame with 16-byte shared-memory loads and TMA copies 16-byte units
  B2S_REQUIRE(plan->shift % 4 == 0, "the fused STFT->PIT kernel needs a shift that is a multiple of 4 (got %d)",
              plan->shift);
  B2S_REQUIRE(frames >= 1, "the fused STFT->PIT kernel needs at least one frame");
  B2S_REQUIRE(samples < ((int64_t)1 << 30) && frames < ((int64_t)1 << 20) && pad_left < ((int64_t)1 << 30),
              "signal too long for the fused STFT->PIT kernel (%lld samples)", (long long)samples);
  constexpr int nt = RECOMPUTE ? K + 1 : K;
  // the hop ring (shift 256, opt-in with B2S_FUSED_RING=1) where its 25 % larger signal area still fits.  Measured on
  // B200 at the north-star shape (profiles/r2_fused_ring.txt): bit-identical results, 11 % fewer shared-memory
  // wavefronts, 26 % less L2 -> SM traffic, 6 % more instructions -- and the same 51.4 us; DRAM traffic is unchanged
  // (139 MB: the overlapping frames of the plain copies hit L2), so it stays off by default.
  constexpr bool kRingFits = sizeof(float) * kFusedWarps * (nt * kRingHops * kHop + 2 * shape.ns * rf::kTile1 + row_area_floats(K)) * shape.ctas
                             + 2048 * shape.ctas <= 228 * 1024;
  const char* re = getenv("B2S_FUSED_RING");
  const bool ring = kRingFits && plan->shift == kHop && pad_left % 4 == 0 && re && atoi(re) != 0;
  const size_t smem = sizeof(float) * kFusedWarps * (nt * (ring ? kRingHops * kHop : rf::kSize) + 2 * shape.ns * rf::kTile1 + row_area_floats(K));
  B2S_REQUIRE(smem <= 227 * 1024, "internal: pipeline shape exceeds the shared memory of an SM");
  auto kernel = ring ? stft_pit_fused_kernel<K, RECOMPUTE, shape.warps, shape.ctas, shape.ns, kRingFits>
                     : stft_pit_fused_kernel<K, RECOMPUTE, shape.warps, shape.ctas, shape.ns, false>;
  static bool configured[2][64] = {};   // per (ring, device) (one static per instantiation)
  if (!configured[ring][plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ring][plan->device & 63] = true;
  }
  static const bool want_trace = getenv("B2S_FUSED_TRACE") != nullptr;
  unsigned long long* trace = nullptr;
  const size_t nstamps = (size_t)g.grid * kFusedWarps * 8;
  (void)smem;
  if (want_trace) {
    B2S_CUDA(cudaMalloc(&trace, nstamps * sizeof(unsigned long long)));
    B2S_CUDA(cudaMemsetAsync(trace, 0, nstamps * sizeof(unsigned long long), stream));
  }
  // Programmatic dependent launch: the kernel may start while its predecessor in the stream (normally the STFT
  // front-end that produces |Y|) drains -- barriers, the constant table and the index arithmetic do not depend on
  // it; the kernel executes griddepcontrol.wait before it requests ANY caller tensor.  B2S_PDL=0 = plain launch.
  static const bool use_pdl = [] { const char* e = getenv("B2S_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(32 * kFusedWarps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  const int shift = plan->shift;
  const float4* table = plan->lane_fwd;
  const int slots = g.slots;
  const char* ab = getenv("B2S_FUSED_ABLATE");   // tuning only (results are wrong with any bit set)
  const int ablate = ab ? atoi(ab) : 0;
  B2S_CUDA(cudaLaunchKernelEx(&cfg, kernel, mixture, yabs, sources, mask, meta, batch, samples, frames, shift,
                              pad_left, table, slots, partial, counters, loss, perm, sse, trace, ablate));
  B2S_LAUNCH_CHECK("stft_pit_fused_kernel");
  if (want_trace) {   // tuning aid: per-warp timeline statistics on stderr
    std::vector<unsigned long long> h(nstamps);
    B2S_CUDA(cudaStreamSynchronize(stream));
    B2S_CUDA(cudaMemcpy(h.data(), trace, nstamps * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    unsigned long long t0 = ~0ull, t_end = 0;
    for (size_t w = 0; w < nstamps / 8; ++w) if (h[w * 8]) { t0 = std::min(t0, h[w * 8]); t_end = std::max(t_end, h[w * 8 + 5]); }
    double sum[6] = {}, mx[6] = {}, mn[6] = {1e30, 1e30, 1e30, 1e30, 1e30, 1e30};
    size_t n = 0;
    for (size_t w = 0; w < nstamps / 8; ++w) {
      if (!h[w * 8] || !h[w * 8 + 5]) continue;
      ++n;
      for (int i = 0; i < 6; ++i) {
        const double v = h[w * 8 + i] ? (double)(h[w * 8 + i] - t0) * 1e-3 : 0.0;
        sum[i] += v; mx[i] = std::max(mx[i], v); mn[i] = std::min(mn[i], v);
      }
    }
    fprintf(stderr, "[fused trace] %zu warps, kernel span %.1f us; event: mean / min / max us after the first warp started\n",
            n, (double)(t_end - t0) * 1e-3);
    const char* names[6] = {"entry", "first copies issued", "position 1 done", "position 2 done", "loop done", "flush done"};
    for (int i = 0; i < 6; ++i) fprintf(stderr, "  %-20s %7.2f / %7.2f / %7.2f\n", names[i], sum[i] / n, mn[i], mx[i]);
    // loop duration (event 4 - event 1) by number of positions of the warp, and per-SM-ish (block) spread
    std::vector<double> dur;
    for (size_t w = 0; w < nstamps / 8; ++w) if (h[w * 8 + 4]) dur.push_back((double)(h[w * 8 + 4] - h[w * 8 + 1]) * 1e-3);
    std::sort(dur.begin(), dur.end());
    if (!dur.empty())
      fprintf(stderr, "  loop duration percentiles 0/10/25/50/75/90/100: %.1f %.1f %.1f %.1f %.1f %.1f %.1f us\n", dur[0],
              dur[dur.size() / 10], dur[dur.size() / 4], dur[dur.size() / 2], dur[dur.size() * 3 / 4],
              dur[dur.size() * 9 / 10], dur.back());
    std::vector<double> fl;
    for (size_t w = 0; w < nstamps / 8; ++w) if (h[w * 8 + 5]) fl.push_back((double)(h[w * 8 + 5] - h[w * 8 + 4]) * 1e-3);
    std::sort(fl.begin(), fl.end());
    if (!fl.empty())
      fprintf(stderr, "  final flush percentiles 0/50/90/99/100: %.2f %.2f %.2f %.2f %.2f us\n", fl[0], fl[fl.size() / 2],
              fl[fl.size() * 9 / 10], fl[fl.size() * 99 / 100], fl.back());
    // mean loop duration of the 4 warps of each block, extremes
    double bmin = 1e30, bmax = 0;
    for (size_t b = 0; b + kFusedWarps <= nstamps / 8; b += kFusedWarps) {
      double s4 = 0;
      for (int w = 0; w < kFusedWarps; ++w) s4 += (double)(h[(b + w) * 8 + 4] - h[(b + w) * 8 + 1]) * 1e-3;
      bmin = std::min(bmin, s4 / kFusedWarps); bmax = std::max(bmax, s4 / kFusedWarps);
    }
    fprintf(stderr, "  per-block mean loop duration: min %.1f max %.1f us\n", bmin, bmax);
    if (const char* path = getenv("B2S_FUSED_TRACE_FILE")) {   // raw stamps: warp, smid, us since the first entry
      if (FILE* f = fopen(path, "w")) {
        for (size_t w = 0; w < nstamps / 8; ++w) {
          if (!h[w * 8]) continue;
          fprintf(f, "%zu %llu", w, h[w * 8 + 6]);
          for (int i = 0; i < 6; ++i) fprintf(f, " %.3f", h[w * 8 + i] ? (double)(h[w * 8 + i] - t0) * 1e-3 : -1.0);
          fprintf(f, "\n");
        }
        fclose(f);
      }
    }
    // loop duration by kind of range: with / without frames that touch the zero padding (first / last 3 frames)
    double se = 0, sn = 0; size_t ne = 0, nn = 0;
    for (int64_t w = 0; w < g.warps; ++w) {
      if (!h[w * 8 + 4]) continue;
      const int64_t pb = w * g.total / g.warps, pe = (w + 1) * g.total / g.warps;
      bool edge = false;
      for (int64_t q = pb; q < pe; ++q) { const int64_t m = q % frames; edge = edge || m < 3 || m >= frames - 3; }
      const double d = (double)(h[w * 8 + 4] - h[w * 8 + 1]) * 1e-3;
      if (edge) { se += d; ++ne; } else { sn += d; ++nn; }
    }
    fprintf(stderr, "  mean loop duration: %zu warps with edge frames %.1f us, %zu without %.1f us\n", ne, ne ? se / ne : 0.0,
            nn, nn ? sn / nn : 0.0);
  }
  return B2S_OK;
}

template <int K>
int launch_fused(const b2s_stft_plan* plan, const float* mixture, const float* yabs, const float* sources,
                 const float* mask, const int64_t* meta, int64_t batch, int64_t samples, int64_t frames,
                 int64_t pad_left, float* loss, int32_t* perm, double* sse, void* workspace,
                 cudaStream_t stream) {
#define B2S_FUSED_ARGS plan, mixture, yabs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, stream
  if (!yabs) return launch_fused_shape<K, true, 0>(B2S_FUSED_ARGS);
  if constexpr (K == 2) {   // two sources: the pair-transform kernel (fused_pair.cuh); B2S_FUSED_PAIR=0 = the 8 x 8 x 8 pipeline
    const char* e = getenv("B2S_FUSED_PAIR");
    const char* w = getenv("B2S_FUSED_WS");
    if (!(e && atoi(e) == 0) && !getenv("B2S_FUSED_VARIANT") && !(w && atoi(w) != 0) && !getenv("B2S_FUSED_RING"))
      return launch_fused_pair(plan, yabs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse,
                               workspace, stream);
  }
  if constexpr (K <= 2) {   // warp-specialised kernel (fused_ws.cuh): opt-in, B2S_FUSED_WS=1 (8 x 8 x 8 transforms) or
    const char* e = getenv("B2S_FUSED_WS");   // =2 (pair transform, K = 2); profiles/r2_fused_experiments.txt
    const int v = e ? atoi(e) : 0;
    if (v == 2 && K == 2)
      return launch_fused_ws<K, K == 2>(plan, yabs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm,
                                        sse, workspace, stream);
    if (v != 0)
      return launch_fused_ws<K, false>(plan, yabs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse,
                                       workspace, stream);
  }
  if (K == 2) {   // tuning alternatives of the headline configuration (tools/hot_bench.py)
    const char* e = getenv("B2S_FUSED_VARIANT");   // read per launch: one process can sweep the shapes
    const int variant = e ? atoi(e) : 0;
    switch (variant) {
      case 1: return launch_fused_shape<K, false, K == 2 ? 1 : 0>(B2S_FUSED_ARGS);
      case 2: return launch_fused_shape<K, false, K == 2 ? 2 : 0>(B2S_FUSED_ARGS);
      case 3: return launch_fused_shape<K, false, K == 2 ? 3 : 0>(B2S_FUSED_ARGS);
      case 4: return launch_fused_shape<K, false, K == 2 ? 4 : 0>(B2S_FUSED_ARGS);
      default: break;
    }
  }
  return launch_fused_shape<K, false, 0>(B2S_FUSED_ARGS);
#undef B2S_FUSED_ARGS
}

}  // namespace

extern "C" {

int64_t b2s_stft_pit_workspace_bytes(int64_t batch, int64_t frames, int sources) {
  if (batch <= 0 || sources <= 0) return kTicketBytes + 16;
  return kTicketBytes + (int64_t)sizeof(double) * batch * fused_max_slots(batch, frames, sources) * sources * sources + 16;
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
  B2S_ON_DEVICE(plan->device);
  cudaStream_t st = (cudaStream_t)stream;
  switch (sources_k) {
    case 1: return launch_fused<1>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
    case 2: return launch_fused<2>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
    case 3: return launch_fused<3>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
    default: return launch_fused<4>(plan, mixture, observation_abs, sources, mask, meta, batch, samples, frames, pad_left, loss, perm, sse, workspace, st);
  }
}

}  // extern "C"
