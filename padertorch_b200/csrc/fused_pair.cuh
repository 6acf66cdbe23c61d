// The fused STFT -> mask -> PIT kernel for TWO sources on the pair transform (cfft_pair.cuh): the two source frames
// of a position go through ONE 1024-point complex FFT (z = s_0 + i s_1), one shared-memory exchange, a mirror
// shuffle and an addition-only separation; lane j then holds bins j + 32 r of BOTH sources, so the K x K SSE reads
// its mask / |Y| operands with unit-stride conflict-free loads and immediates (no bin tables, no lane-0 cases
// except bin 512).  Included by fused.cu (inside its anonymous namespace, after the shared helpers); everything
// around the transform -- ranges of frame positions per warp, TMA-fed frames and rows, partial-sum slots, ticket,
// fold, permutation search, programmatic launch -- is that of stft_pit_fused_kernel.
#pragma once
// (cfft_pair.cuh is included by fused.cu at namespace scope)

#ifndef B2S_PAIR_EARLY_ROWS
#define B2S_PAIR_EARLY_ROWS 0
#endif
#ifndef B2S_PAIR_PACKED_WINDOW
#define B2S_PAIR_PACKED_WINDOW 0
#endif
constexpr int kPairWarps = 4, kPairCtas = 2;
// RING (shift 256): the source frames are not copied whole.  Each source keeps a ring of five hops of 256 samples; the
// walk visits consecutive frames of an utterance, so frame m + 1 needs ONE new hop per source (1 KB instead of 4 KB),
// and that hop's slot is not part of the frame being transformed: it is requested at the START of position m -- a
// whole position ahead instead of two thirds of one -- on the other of two mbarriers.
constexpr int kPairRingHops = 5;
__host__ __device__ constexpr int pair_sig_floats(bool ring) { return ring ? kPairRingHops * kHop : rf::kSize; }
__host__ __device__ constexpr int pair_warp_floats(bool ring) {
  return 2 * pair_sig_floats(ring) + 2 * 32 * cp::kPitch + row_area_floats(2);   // two sources, the transpose tile, rows
}

template <bool RING>
__global__ void __launch_bounds__(32 * kPairWarps, kPairCtas)
stft_pit_pair_kernel(const float* __restrict__ yabs, const float* __restrict__ sources, const float* __restrict__ mask,
                     const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames, int shift,
                     int64_t pad_left, const float* __restrict__ window, const float2* __restrict__ tab, int slots,
                     double* __restrict__ partial, int* __restrict__ counters, float* __restrict__ loss,
                     int32_t* __restrict__ perm, double* __restrict__ sse) {
  constexpr int K = 2, NV = 4;
  constexpr int F = rf::kBins;
  constexpr int kSig = pair_sig_floats(RING);
  constexpr int kWarpFloats = pair_warp_floats(RING);
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t bars[kPairWarps][3];   // frames (two, alternating by position with RING), rows
  __shared__ double totals_sm[kPairWarps][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sig = smem + warp * kWarpFloats;                               // source t at sig + t * kSig
  float2* tile = reinterpret_cast<float2*>(sig + 2 * kSig);             // [32][kPitch]
  float* rows_area = sig + 2 * kSig + 2 * 32 * cp::kPitch;
  uint64_t* bar_sig = &bars[warp][0];
  uint64_t* bar_rows = &bars[warp][2];
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    mbar_init(bar_rows, 1);
    fence_mbar_init();
  }
  __syncwarp();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // per-lane constants from the plan's tables (not caller data: loaded before the dependency wait)
  cp::PairConsts k;
  k.lane = lane;
#pragma unroll
  for (int p = 0; p < 32; ++p) k.w[p] = __ldg(window + lane + 32 * p);   // plan->awin_half: no arithmetic behind the loads
#pragma unroll
  for (int q = 0; q < 32; ++q) k.t[cp::out_pos(q)] = __ldg(tab + 32 * q + lane)   /* plan->pair_tw: [q][lane] */;

  const int64_t total = batch * frames;
  const int64_t nwarps = min((int64_t)gridDim.x * kPairWarps, total);
  const int64_t gw = (int64_t)blockIdx.x * kPairWarps + warp;
  if (gw >= nwarps) return;
  const int64_t p_begin = range_start(gw, total, nwarps), p_end = range_start(gw + 1, total, nwarps);

  unsigned sig_phase = 0, rows_phase = 0;
  bool sig_by_tma = false;
  int off_m = 0, off_y = 0;
  auto frames_of = [&](int64_t b) { return meta ? meta[2 * b + 1] : frames; };
  // The context (row pointers, lengths) is that of the example whose position is REQUESTED next; it changes only when
  // the walk below crosses an example boundary, so the steady state has no per-position context checks.
  int ctx_T = 0, ctx_M = 0;
  bool ctx_a16 = false;
  const float* ctx_row[K];
  const float* ctx_mask = nullptr;
  const float* ctx_y = nullptr;
  auto load_ctx = [&](int64_t b) {
    ctx_T = (int)(meta ? meta[2 * b] : samples);
    ctx_M = (int)frames_of(b);
    ctx_a16 = true;
#pragma unroll
    for (int t = 0; t < K; ++t) {
      ctx_row[t] = sources + (b * K + t) * samples;
      ctx_a16 = ctx_a16 && (reinterpret_cast<uintptr_t>(ctx_row[t]) & 15) == 0;
    }
    ctx_mask = mask + b * frames * (K * F);
    ctx_y = yabs + b * frames * F;
  };
  const int pad = (int)pad_left;
  // request the two source frames of frame m of the context's example: TMA, or zero-filling cp.async for frames that
  // touch the zero padding at the signal's ends or are not 16-byte aligned
  auto start_signals = [&](int m) {
    const int s0 = m * shift - pad;
    const bool a16 = ctx_a16 && (s0 & 3) == 0;
    const bool bulk = a16 && s0 >= 0 && s0 + rf::kSize <= ctx_T;
    sig_by_tma = bulk;
    if (bulk) {
      if (lane == 0) {
        mbar_expect_tx(bar_sig, K * rf::kSize * 4u);
#pragma unroll
        for (int t = 0; t < K; ++t) bulk_g2s(sig + t * rf::kSize, ctx_row[t] + s0, rf::kSize * 4u, bar_sig);
      }
    } else {
#pragma unroll
      for (int t = 0; t < K; ++t) {
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
  // RING: request hops [h0, h1) of the context's example on frame barrier `which`; bit `which` of the two masks says
  // how they travel (bulk copies completing on the barrier / zero-filling cp.async for hops that touch the padding at
  // the signal's ends or are not 16-byte aligned)
  unsigned ring_phase = 0, ring_bulk = 0, ring_async = 0;
  auto request_hops = [&](int h0, int h1, int which) {
    int nbulk = 0;
    for (int h = h0; h < h1; ++h) {
      const int s0 = h * kHop - pad;
      nbulk += (ctx_a16 && (s0 & 3) == 0 && s0 >= 0 && s0 + kHop <= ctx_T) ? 1 : 0;
    }
    const bool any_async = nbulk < h1 - h0;
    ring_bulk = nbulk > 0 ? (ring_bulk | (1u << which)) : (ring_bulk & ~(1u << which));
    ring_async = any_async ? (ring_async | (1u << which)) : (ring_async & ~(1u << which));
    if (nbulk > 0 && lane == 0) mbar_expect_tx(&bars[warp][which], (unsigned)(nbulk * K * kHop * 4));
    for (int h = h0; h < h1; ++h) {
      const int s0 = h * kHop - pad;
      float* dst = sig + (h % kPairRingHops) * kHop;
      if (ctx_a16 && (s0 & 3) == 0 && s0 >= 0 && s0 + kHop <= ctx_T) {
        if (lane == 0) {
#pragma unroll
          for (int t = 0; t < K; ++t) bulk_g2s(dst + t * kSig, ctx_row[t] + s0, kHop * 4u, &bars[warp][which]);
        }
      } else {
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const float* xr = ctx_row[t];
          for (int i = lane; i < kHop; i += 32) {
            const int n = s0 + i;
            const bool ok = n >= 0 && n < ctx_T;
            fft::cp_async_4_zfill(dst + t * kSig + i, ok ? xr + n : xr, ok ? 4 : 0);
          }
        }
      }
    }
    if (any_async) fft::cp_async_commit();
  };
  // request the mask rows [K][F] and the |Y| row of frame m of the context's example (enclosing 16-byte aligned
  // ranges; the rows sit at the source's misalignment inside the landing areas)
  auto start_rows = [&](int m) {
    const uintptr_t am = reinterpret_cast<uintptr_t>(ctx_mask + m * (K * F));
    const uintptr_t ay = reinterpret_cast<uintptr_t>(ctx_y + m * F);
    off_m = (int)(am & 15) >> 2;
    off_y = (int)(ay & 15) >> 2;
    if (lane == 0) {
      const unsigned bytes_m = (unsigned)(((am & 15) + K * F * 4 + 15) & ~15u);
      const unsigned bytes_y = (unsigned)(((ay & 15) + F * 4 + 15) & ~15u);
      mbar_expect_tx(bar_rows, bytes_m + bytes_y);
      bulk_g2s(rows_area, reinterpret_cast<const void*>(am & ~(uintptr_t)15), bytes_m, bar_rows);
      bulk_g2s(rows_area + mask_area_floats(K), reinterpret_cast<const void*>(ay & ~(uintptr_t)15), bytes_y, bar_rows);
    }
  };

  float2 acc[NV];   // two bins per packed accumulator
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float2(0.f, 0.f);
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
      constexpr int per = 32 / NV;
      const int vi = lane % NV, c0 = lane / NV;
      double s = 0.0;
      {
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

  const int frames_i = (int)frames;
  const int partner = (32 - lane) & 31;
  const bool first = lane == 0;
  // The walk over the LIVE positions of the range (frames beyond an example's own length -- ragged batches -- are
  // skipped): (q, b, m) is the position being processed, (qn, bn, mn) the next live one (qn == p_end: none).  In the
  // steady state the next position is the next frame of the same example: one comparison.
  // Every warp whose (dense) range touches an example contributes a partial sum and a ticket for it, live frames or
  // not: examples that the walk passes without a live position are flushed with zero sums.
  const int64_t b_last = (p_end - 1) / frames;
  int64_t q = p_begin, b = p_begin / frames;
  int m = (int)(p_begin - b * frames);
  asm volatile("griddepcontrol.wait;" ::: "memory");   // nothing of the caller's tensors is requested before this
  load_ctx(b);
  while (m >= ctx_M) {   // the range starts in the padding of a shorter example
    flush(b);
    q += frames_i - m;
    ++b; m = 0;
    if (q >= p_end) return;
    load_ctx(b);
  }
  int which = 0;   // RING: frame barrier of the current position
  if (RING) request_hops(m, m + 4, 0); else start_signals(m);
  start_rows(m);
  for (;;) {
    int64_t qn = q + 1, bn = b;
    int mn = m + 1;
    const bool same = qn < p_end && mn < ctx_M;   // next frame of the same example: the context stays
    // RING: the next frame of the same example adds hop m + 4, whose slot the current frame does not use: request it now
    if (RING && same) request_hops(m + 4, m + 5, which ^ 1);
    auto find_next = [&]() {            // rare: example boundary, padding frames, end of the range
      for (;;) {
        if (qn >= p_end) { qn = p_end; bn = b_last + 1; return; }
        if (mn >= frames_i) { ++bn; mn = 0; }
        if (bn != b) load_ctx(bn);       // (for bn == b the context is already that example's)
        if (mn < ctx_M) return;
        qn += frames_i - mn;             // the rest of this example is padding
        mn = frames_i;
      }
    };
    if (RING) {
      if (ring_bulk & (1u << which)) {
        mbar_wait(&bars[warp][which], (ring_phase >> which) & 1u);
        ring_phase ^= 1u << which;
      }
      if (ring_async & (1u << which)) {   // (also waits for the next position's hop if that one travels by cp.async)
        fft::cp_async_wait_all();
        __syncwarp();
      }
    } else if (sig_by_tma) {
      mbar_wait(bar_sig, sig_phase);
      sig_phase ^= 1;
    } else {
      fft::cp_async_wait_all();
      __syncwarp();
    }
    // ---- pass 1: z[n] = w[n] (s_0[n] + i s_1[n]), n = lane + 32 p: conflict-free strided loads, radix-32 in registers
    float2 v[32];
    if (RING) {   // sample lane + 32 p lies in hop p / 8 of the frame: ring slot (m + p / 8) % 5
      const int base = m % kPairRingHops;
      const float* fh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) fh[j] = sig + lane + ((base + j >= kPairRingHops) ? base + j - kPairRingHops : base + j) * kHop;
#pragma unroll
      for (int p = 0; p < 32; ++p)
        v[p] = make_float2(k.w[p] * fh[p / 8][32 * (p % 8)], k.w[p] * fh[p / 8][kSig + 32 * (p % 8)]);
    } else {
      const float* fa = sig + lane;
      const float* fb = sig + rf::kSize + lane;
#pragma unroll
      for (int p = 0; p < 32; ++p) v[p] = make_float2(k.w[p] * fa[32 * p], k.w[p] * fb[32 * p]);
    }
    __syncwarp();                      // every lane holds its samples: the frames may be overwritten
    if (!same) find_next();            // (switches the context to the next position's example)
    if (qn < p_end) {
      // RING: a continuation was requested at the top; the first frame of another example fills four hops (the ring's
      // slots are free: every lane has its samples)
      if (!RING) start_signals(mn); else if (!same) request_hops(mn, mn + 4, which ^ 1);
    }
    cp::radix32(v);
#pragma unroll
    for (int qq = 0; qq < 32; ++qq) {
      const float2 u = v[cp::out_pos(qq)];
      tile[lane * cp::kPitch + qq] = qq == 0 ? u : rf::cmul(u, k.t[cp::out_pos(qq)]);
    }
    __syncwarp();
    // ---- pass 2: column `lane` of the transpose, radix-32; Z[lane + 32 r] = v[r]
#pragma unroll
    for (int l = 0; l < 32; ++l) v[l] = tile[l * cp::kPitch + lane];
    __syncwarp();                      // the tile may be rewritten by the next position
    cp::radix32(v);
#if B2S_PAIR_EARLY_ROWS
    // the rows of this position were requested a whole transform ago: waiting for them here puts their loads into the
    // same straight-line block as the shuffles, square roots and the SSE
    mbar_wait(bar_rows, rows_phase);
    rows_phase ^= 1;
#endif
    // ---- mirror exchange with lane (32 - lane) % 32, separation, magnitudes: bins lane + 32 r of both sources
    float xa[17], xb[17];              // |STFT(s_0)|, |STFT(s_1)|; slot 16 = bin 512 (lane 0)
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 z = v[cp::out_pos(r)];
      const float2 hi = v[cp::out_pos(31 - r)], lo = v[cp::out_pos((32 - r) & 31)];
      const float2 send = first ? lo : hi;
      float2 mm;
      mm.x = __shfl_sync(0xffffffffu, send.x, partner);
      mm.y = __shfl_sync(0xffffffffu, send.y, partner);
      const float2 cm = make_float2(mm.x, -mm.y);
      const float2 s = rf::add2(z, cm), d = rf::sub2(z, cm);     // A = s, B = -i d (the window carries the 1/2)
      xa[r] = fft::sqrt_approx(fmaf(s.x, s.x, s.y * s.y));
      xb[r] = fft::sqrt_approx(fmaf(d.x, d.x, d.y * d.y));
    }
    {
      const float2 zn = v[cp::out_pos(16)];   // lane 0: Z[512] = (A[512] + i B[512]) / 2, both real
      xa[16] = 2.f * fabsf(zn.x);
      xb[16] = 2.f * fabsf(zn.y);
    }
    // ---- SSE of this frame: e_i = mask_i * |Y| against both source magnitudes
#if !B2S_PAIR_EARLY_ROWS
    mbar_wait(bar_rows, rows_phase);
    rows_phase ^= 1;
#endif
    const float* mrow = rows_area + off_m + lane;
    const float* yrow = rows_area + mask_area_floats(K) + off_y + lane;
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
      const float2 ov = make_float2(yrow[32 * r], yrow[32 * r + 32]);
      const float2 x0 = make_float2(xa[r], xa[r + 1]), x1 = make_float2(xb[r], xb[r + 1]);
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float2 e = rf::mul2(make_float2(mrow[i * F + 32 * r], mrow[i * F + 32 * r + 32]), ov);
        const float2 d0 = rf::sub2(e, x0), d1 = rf::sub2(e, x1);
        acc[i * K + 0] = rf::fma2(d0, d0, acc[i * K + 0]);
        acc[i * K + 1] = rf::fma2(d1, d1, acc[i * K + 1]);
      }
    }
    {   // bin 512: every lane reads the same words, lane 0's squares count
      const float oy = yrow[rf::kHalf - lane];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float e = mrow[i * F + rf::kHalf - lane] * oy;
        const float d0 = e - xa[16], d1 = e - xb[16];
        acc[i * K + 0].x = fmaf(first ? d0 : 0.f, d0, acc[i * K + 0].x);
        acc[i * K + 1].x = fmaf(first ? d1 : 0.f, d1, acc[i * K + 1].x);
      }
    }
    __syncwarp();   // every lane has read its rows: the area may be overwritten
    if (qn < p_end) start_rows(mn);
    if (bn != b) {   // warp-uniform: the range leaves example b (and possibly passes examples without a live frame)
      for (int64_t bb = b; bb < bn; ++bb) flush(bb);
      if (qn >= p_end) break;
    }
    q = qn; b = bn; m = mn;
    which ^= 1;
  }
}

int launch_fused_pair(const b2s_stft_plan* plan, const float* yabs, const float* sources, const float* mask,
                      const int64_t* meta, int64_t batch, int64_t samples, int64_t frames, int64_t pad_left,
                      float* loss, int32_t* perm, double* sse, void* workspace, cudaStream_t stream) {
  B2S_REQUIRE(plan->shift % 4 == 0, "the fused STFT->PIT kernel needs a shift that is a multiple of 4 (got %d)",
              plan->shift);
  B2S_REQUIRE(frames >= 1, "the fused STFT->PIT kernel needs at least one frame");
  B2S_REQUIRE(samples < ((int64_t)1 << 30) && frames < ((int64_t)1 << 20) && pad_left < ((int64_t)1 << 30),
              "signal too long for the fused STFT->PIT kernel (%lld samples)", (long long)samples);
  const FusedGrid g = fused_grid(batch, frames, FusedShape{kPairWarps, kPairCtas, 1});
  // the hop ring (shift 256, 16-byte granular padding) is opt-in, B2S_PAIR_RING=1: measured 45.3 us against 44.0 us
  // with whole frames per position (profiles/r2_fused_experiments.txt) -- a quarter of the L2 -> shared-memory bytes
  // and a whole position of prefetch distance do not pay for the slot arithmetic
  const char* re = getenv("B2S_PAIR_RING");
  const bool ring = plan->shift == kHop && pad_left % 4 == 0 && re && atoi(re) != 0;
  const size_t smem = sizeof(float) * kPairWarps * pair_warp_floats(ring);
  static_assert(sizeof(float) * kPairWarps * pair_warp_floats(true) * kPairCtas + 2048 <= 227 * 1024,
                "pipelines exceed the shared memory of an SM");
  auto kernel = ring ? stft_pit_pair_kernel<true> : stft_pit_pair_kernel<false>;
  static bool configured[2][64] = {};
  if (!configured[ring][plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ring][plan->device & 63] = true;
  }
  static const bool use_pdl = [] { const char* e = getenv("B2S_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(32 * kPairWarps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  const int shift = plan->shift;
  const float* window = plan->awin_half;
  const float2* tab = plan->pair_tw;
  const int slots = g.slots;
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  B2S_CUDA(cudaLaunchKernelEx(&cfg, kernel, yabs, sources, mask, meta, batch, samples, frames, shift,
                              pad_left, window, tab, slots, partial, counters, loss, perm, sse));
  B2S_LAUNCH_CHECK("stft_pit_pair_kernel");
  return B2S_OK;
}
