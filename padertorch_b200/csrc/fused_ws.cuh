// Warp-specialised form of the fused STFT -> mask -> PIT kernel (K <= 2, |Y| read from the front-end's output).
// Included by fused.cu (inside its anonymous namespace, after the shared helpers).
//
// OPT-IN (B2S_FUSED_WS=1): measured SLOWER than the one-role kernel on B200 -- 54.2 us against 47.6 us at the
// north-star shape (profiles/r2_fused_experiments.txt): the magnitudes cross shared memory (76 more wavefronts per
// position on a pipe that is already ~70 % busy), which costs more than the serial SSE it removes from the transform
// warps.  Kept as the measured record of that design (parity-green: tests/test_gpu_round2.py) and as the working
// example of register-file re-splitting with setmaxnreg.  compute-sanitizer: memcheck clean; racecheck reports two
// WARNINGS for this kernel only (frame-buffer refill against the same warp's pass-1 loads -- the sequence the tool
// accepts in the one-role kernels, profiles/r2_sanitizer.txt).
//
// Idea: the one-role pipeline of stft_pit_fused_kernel executes transform, magnitudes, row copies, K x K SSE and the
// example bookkeeping serially in every warp, at two resident warps per SM sub-partition (240 registers).  Without the
// SSE and the row copies the same loop needs 13 % less time -- independent work which the register budget keeps from
// running next to the transforms.  Here it runs in its own warps:
//   warps 0..3   (one warp group, 120 registers after setmaxnreg.dec)  SSE warps: each serves two streams -- requests
//                the mask rows [K][513] and the |Y| row of a position by TMA, waits for the position's magnitudes,
//                accumulates the K x K SSE (lane j owns bins j + 32 r: every shared-memory operand is a conflict-free
//                unit-stride LDS.32, no bin tables, no lane-0 special cases), flushes partial sums at example
//                boundaries and runs the ticket / fold / permutation search;
//   warps 4..11  (two warp groups, 192 registers after setmaxnreg.inc)  transform warps = streams: TMA-fed frames of
//                the K sources -> rfft_streams -> magnitudes -> [K][513] floats in shared memory (conflict-free STS.32)
//                -> mbarrier `full`; the buffer is handed back by the SSE warp through mbarrier `empty`.
// One CTA of 384 threads per SM (launched with 168 registers per thread = the whole register file; the two roles then
// re-split it).  Stream ranges, partial-sum slots, tie-break and results' layout are those of stft_pit_fused_kernel.
#pragma once

constexpr int kWsStreams = 8;                                   // transform warps per CTA
constexpr int kWsSseWarps = 4;                                  // one warp group
constexpr int kWsThreads = 32 * (kWsStreams + kWsSseWarps);     // 384
constexpr int kWsMagRow = 520;                                  // floats per magnitude row (513 bins, 16-byte multiple)
constexpr int kWsSseRegs = 120, kWsTransformRegs = 192;         // 128 * 120 + 256 * 192 = 384 * 168
constexpr int kWsSseRegsPair = 104, kWsTransformRegsPair = 200;  // 128 * 104 + 256 * 200 = 384 * 168

__host__ __device__ constexpr int ws_stream_floats(int K) {
  return K * rf::kSize + 2 * 2 * rf::kTile1 + K * kWsMagRow + row_area_floats(K);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// PAIR (K == 2): the transform warps run the pair transform (cfft_pair.cuh) instead of two 8 x 8 x 8 transforms.
template <int K, bool PAIR>
__global__ void __launch_bounds__(kWsThreads, 1)
stft_pit_ws_kernel(const float* __restrict__ yabs, const float* __restrict__ sources, const float* __restrict__ mask,
                   const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames, int shift,
                   int64_t pad_left, const float4* __restrict__ lane_table, int slots, double* __restrict__ partial,
                   int* __restrict__ counters, float* __restrict__ loss, int32_t* __restrict__ perm,
                   double* __restrict__ sse, const float* __restrict__ window, const float2* __restrict__ tab,
                   int exp_arg /* tuning builds only: 1 SSE warps skip the arithmetic, 2 no SSE warps */) {
  static_assert(!PAIR || K == 2, "the pair transform takes the two sources of a position");
  const int exp = kTuning ? exp_arg : 0;
  constexpr int NV = K * K;
  constexpr int F = rf::kBins;
  constexpr int kStreamFloats = ws_stream_floats(K);
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t bars[kWsStreams][4];   // frames landed, rows landed, magnitudes full, magnitudes empty
  __shared__ double totals_sm[kWsSseWarps][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kWsStreams; ++s) {
      mbar_init(&bars[s][0], 1);
      mbar_init(&bars[s][1], 1);
      mbar_init(&bars[s][2], 32);
      mbar_init(&bars[s][3], 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int64_t total = batch * frames;
  const int64_t nstreams = min((int64_t)gridDim.x * kWsStreams, total);   // every stream owns >= 1 position
  const int frames_i = (int)frames;
  auto frames_of = [&](int64_t b) { return meta ? meta[2 * b + 1] : frames; };

  if (warp >= kWsSseWarps) {
    // ================================================= transform warps ==========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PAIR ? kWsTransformRegsPair : kWsTransformRegs));
    const int st = warp - kWsSseWarps;
    const int64_t gs = (int64_t)blockIdx.x * kWsStreams + st;
    if (gs >= nstreams) return;
    float* sig = smem + st * kStreamFloats;                       // frame of source t at sig + t * 1024
    float2* tile = reinterpret_cast<float2*>(sig + K * rf::kSize);
    float* mag = sig + K * rf::kSize + 2 * 2 * rf::kTile1;        // [K][kWsMagRow]
    uint64_t* bar_sig = &bars[st][0];
    uint64_t* bar_full = &bars[st][2];
    uint64_t* bar_empty = &bars[st][3];
    rf::LaneConsts k;
    cp::PairConsts kp;
    if (PAIR) {
      kp.lane = lane;
#pragma unroll
      for (int p = 0; p < 32; ++p) kp.w[p] = 0.5f * __ldg(window + lane + 32 * p);
#pragma unroll
      for (int q = 0; q < 32; ++q) kp.t[cp::out_pos(q)] = __ldg(tab + 32 * q + lane)   /* plan->pair_tw: [q][lane] */;
    } else {
      k.load(lane_table, lane);
    }
    // slot p holds bin (p < 4 ? k0 : k4) + 64 p on the A side and 512 minus that on the B side
    const int k0 = rf::bin_a(lane, 0), k4 = rf::bin_a(lane, 4) - 256;
    const bool first = lane == 0;
    const int64_t p_begin = range_start(gs, total, nstreams), p_end = range_start(gs + 1, total, nstreams);
    const int pad = (int)pad_left;

    unsigned sig_phase = 0, empty_phase = 1;   // a fresh barrier passes a wait on parity 1: the buffer starts empty
    bool sig_by_tma = false;
    int64_t ctx_b = -1;
    int ctx_T = 0, ctx_M = 0;
    bool ctx_a16 = false;
    const float* ctx_row[K];
    auto set_ctx = [&](int64_t b) {
      if (b == ctx_b) return;
      ctx_b = b;
      ctx_T = (int)(meta ? meta[2 * b] : samples);
      ctx_M = (int)frames_of(b);
      ctx_a16 = true;
#pragma unroll
      for (int t = 0; t < K; ++t) {
        ctx_row[t] = sources + (b * K + t) * samples;
        ctx_a16 = ctx_a16 && (reinterpret_cast<uintptr_t>(ctx_row[t]) & 15) == 0;
      }
    };
    // start the copy of the K frames of position q = (b, m): TMA, or zero-filling cp.async for frames that touch the
    // zero padding at the signal's ends or are not 16-byte aligned
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
    // magnitudes of one transform -> its row in shared memory (bins in natural order)
    auto store_magnitudes = [&](float* row, const float2 (&ya)[8], const float2 (&yb)[8], float ydc, float ynyq) {
      float* pa = row + k0;
      float* pb = row + (rf::kHalf - k0);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float2 x = mag2(ya[p], yb[p]);
        const int o = (p < 4 ? 0 : k4 - k0) + 64 * p;   // lanes >= 1: k4 == k0
        pa[o] = x.x;
        pb[-o] = x.y;                                    // lane 0, slot 7: bin 256 on both sides (same value twice)
      }
      if (first) {
        row[0] = fabsf(ydc);
        row[rf::kHalf] = fabsf(ynyq);
      }
    };

    int64_t b = p_begin / frames;
    int m = (int)(p_begin - b * frames);
    asm volatile("griddepcontrol.wait;" ::: "memory");   // nothing of the caller's tensors is requested before this
    start_signals(p_begin, b, m);
    for (int64_t q = p_begin; q < p_end; ++q) {
      int64_t bn = b;
      int mn = m + 1;
      if (mn == frames_i) { mn = 0; ++bn; }
      set_ctx(b);
      if (m >= ctx_M) {   // position beyond this example's length (ragged batch)
        start_signals(q + 1, bn, mn);
        b = bn; m = mn;
        continue;
      }
      if (sig_by_tma) {
        mbar_wait(bar_sig, sig_phase);
        sig_phase ^= 1;
      } else {
        fft::cp_async_wait_all();
        __syncwarp();
      }
      auto next_copy = [&]() { start_signals(q + 1, bn, mn); };
      if (PAIR) {
        float2 v[32];
        {
          const float* fa = sig + lane;
          const float* fb = sig + rf::kSize + lane;
#pragma unroll
          for (int p = 0; p < 32; ++p) v[p] = make_float2(kp.w[p] * fa[32 * p], kp.w[p] * fb[32 * p]);
        }
        __syncwarp();   // every lane holds its samples (and the previous position's tile reads are done)
        next_copy();
        cp::radix32(v);
#pragma unroll
        for (int qq = 0; qq < 32; ++qq) {
          const float2 u = v[cp::out_pos(qq)];
          tile[lane * cp::kPitch + qq] = qq == 0 ? u : rf::cmul(u, kp.t[cp::out_pos(qq)]);
        }
        __syncwarp();
#pragma unroll
        for (int l = 0; l < 32; ++l) v[l] = tile[l * cp::kPitch + lane];
        cp::radix32(v);
        const int partner = (32 - lane) & 31;
        float xa[16], xb[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float2 z = v[cp::out_pos(r)];
          const float2 hi = v[cp::out_pos(31 - r)], lo = v[cp::out_pos((32 - r) & 31)];
          const float2 send = first ? lo : hi;
          float2 mm;
          mm.x = __shfl_sync(0xffffffffu, send.x, partner);
          mm.y = __shfl_sync(0xffffffffu, send.y, partner);
          const float2 cm = make_float2(mm.x, -mm.y);
          const float2 sa = rf::add2(z, cm), d = rf::sub2(z, cm);
          xa[r] = fft::sqrt_approx(fmaf(sa.x, sa.x, sa.y * sa.y));
          xb[r] = fft::sqrt_approx(fmaf(d.x, d.x, d.y * d.y));
        }
        const float2 zn = v[cp::out_pos(16)];
        if (!(exp & 2)) mbar_wait(bar_empty, empty_phase);   // the SSE warp has read the previous position's magnitudes
        empty_phase ^= 1;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          mag[lane + 32 * r] = xa[r];
          mag[kWsMagRow + lane + 32 * r] = xb[r];
        }
        if (first) {
          mag[rf::kHalf] = 2.f * fabsf(zn.x);
          mag[kWsMagRow + rf::kHalf] = 2.f * fabsf(zn.y);
        }
      } else if (K == 2) {
        float2 ya[2][8], yb[2][8];
        float ydc[2], ynyq[2];
        rf::rfft_streams<2, false, false>(sig, rf::kSize, tile, k, ya, yb, ydc, ynyq, 0, next_copy);
        if (!(exp & 2)) mbar_wait(bar_empty, empty_phase);   // the SSE warp has read the previous position's magnitudes
        empty_phase ^= 1;
        store_magnitudes(mag, ya[0], yb[0], ydc[0], ynyq[0]);
        store_magnitudes(mag + kWsMagRow, ya[1], yb[1], ydc[1], ynyq[1]);
      } else {
        float2 ya[1][8], yb[1][8];
        float ydc[1], ynyq[1];
        rf::rfft_streams<1, false, false>(sig, 0, tile, k, ya, yb, ydc, ynyq, 0, next_copy);
        mbar_wait(bar_empty, empty_phase);
        empty_phase ^= 1;
        store_magnitudes(mag, ya[0], yb[0], ydc[0], ynyq[0]);
      }
      mbar_arrive(bar_full);   // every lane after its own stores (release): 32 arrivals complete the phase
      b = bn; m = mn;
    }
    return;
  }

  // ===================================================== SSE warps ================================================
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PAIR ? kWsSseRegsPair : kWsSseRegs));
  if (exp & 2) return;
  constexpr int U = kWsStreams / kWsSseWarps;   // streams per SSE warp
  int64_t q[U], p_end[U], gsu[U];
  int64_t b[U], b_cur[U], ctx_b[U];
  int m[U], ctx_M[U], off_m[U], off_y[U];
  unsigned rows_phase[U], full_phase[U];
  const float* ctx_mask[U];
  const float* ctx_y[U];
  float2 acc[U][NV];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    gsu[u] = (int64_t)blockIdx.x * kWsStreams + warp * U + u;
    const bool active = gsu[u] < nstreams;
    q[u] = active ? range_start(gsu[u], total, nstreams) : 0;
    p_end[u] = active ? range_start(gsu[u] + 1, total, nstreams) : 0;
    b[u] = q[u] / frames;
    m[u] = (int)(q[u] - b[u] * frames);
    b_cur[u] = active ? b[u] : -1;
    ctx_b[u] = -1;
    ctx_M[u] = 0; off_m[u] = 0; off_y[u] = 0;
    rows_phase[u] = 0; full_phase[u] = 0;
    ctx_mask[u] = nullptr; ctx_y[u] = nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[u][i] = make_float2(0.f, 0.f);
  }
  auto stream_base = [&](int u) { return smem + (warp * U + u) * kStreamFloats; };
  auto set_ctx = [&](int u, int64_t bb) {
    if (bb == ctx_b[u]) return;
    ctx_b[u] = bb;
    ctx_M[u] = (int)frames_of(bb);
    ctx_mask[u] = mask + bb * frames * (K * F);
    ctx_y[u] = yabs + bb * frames * F;
  };
  // Start the copy of the mask rows [K][F] (contiguous) and the |Y| row of position qq of stream u: the enclosing
  // 16-byte aligned ranges are copied, the rows sit at the source's misalignment inside the landing areas (the bytes
  // before / after a row belong to the neighbouring rows or to the same >= 256-byte granular allocation).
  auto start_rows = [&](int u, int64_t qq, int64_t bb, int mm) {
    if (qq >= p_end[u]) return;
    set_ctx(u, bb);
    if (mm >= ctx_M[u]) return;
    const uintptr_t am = reinterpret_cast<uintptr_t>(ctx_mask[u] + mm * (K * F));
    const uintptr_t ay = reinterpret_cast<uintptr_t>(ctx_y[u] + mm * F);
    off_m[u] = (int)(am & 15) >> 2;
    off_y[u] = (int)(ay & 15) >> 2;
    if (lane == 0) {
      float* rows_area = stream_base(u) + K * rf::kSize + 2 * 2 * rf::kTile1 + K * kWsMagRow;
      uint64_t* bar_rows = &bars[warp * U + u][1];
      const unsigned bytes_m = (unsigned)(((am & 15) + K * F * 4 + 15) & ~15u);
      const unsigned bytes_y = (unsigned)(((ay & 15) + F * 4 + 15) & ~15u);
      mbar_expect_tx(bar_rows, bytes_m + bytes_y);
      bulk_g2s(rows_area, reinterpret_cast<const void*>(am & ~(uintptr_t)15), bytes_m, bar_rows);
      bulk_g2s(rows_area + mask_area_floats(K), reinterpret_cast<const void*>(ay & ~(uintptr_t)15), bytes_y, bar_rows);
    }
  };
  // flush the partial sums of example bb of stream u (whole warp); the last stream of an example (ticket) folds the
  // partials in slot order and searches the K! permutations
  auto flush = [&](int u, int64_t bb) {
    double mine[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      mine[i] = (double)warp_sum(acc[u][i].x + acc[u][i].y);
      acc[u][i] = make_float2(0.f, 0.f);
    }
    const int64_t first_owner = owner_of(bb * frames, total, nstreams);
    const int64_t last_owner = owner_of((bb + 1) * frames - 1, total, nstreams);
    const int slot = (int)(gsu[u] - first_owner), nparts = (int)(last_owner - first_owner + 1);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) partial[(bb * slots + slot) * NV + i] = mine[i];
      __threadfence();
    }
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(counters + bb, 1) == nparts - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {   // warp-uniform
      __threadfence();
      double* totals = totals_sm[warp];
      constexpr int per = 32 / NV;
      const int vi = lane % NV, c0 = lane / NV;
      double s = 0.0;
      if (c0 < per) {
        const volatile double* p = partial + bb * slots * NV + vi;
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
        sse[bb * NV + lane] = s;
      }
      __syncwarp();
      double best;
      int bp[B2S_MAX_SOURCES];
      warp_search_permutations(totals, K, lane, best, bp);
      if (lane == 0) {
        loss[bb] = (float)(best / ((double)frames_of(bb) * (double)K * (double)F));
        for (int kk = 0; kk < K; ++kk) perm[bb * K + kk] = bp[kk];
        counters[bb] = 0;
      }
      __syncwarp();
    }
  };
  // one position of stream u
  auto step = [&](int u) {
    int64_t bn = b[u];
    int mn = m[u] + 1;
    if (mn == frames_i) { mn = 0; ++bn; }
    if (b[u] != b_cur[u]) {   // warp-uniform
      flush(u, b_cur[u]);
      b_cur[u] = b[u];
    }
    set_ctx(u, b[u]);
    if (m[u] < ctx_M[u]) {
      const float* base = stream_base(u);
      const float* mg = base + K * rf::kSize + 2 * 2 * rf::kTile1 + lane;
      const float* mrow = mg + K * kWsMagRow + off_m[u];
      const float* yrow = mg + K * kWsMagRow + mask_area_floats(K) + off_y[u];
      uint64_t* bs = bars[warp * U + u];
      mbar_wait(&bs[1], rows_phase[u]);   // the rows of this position have landed
      rows_phase[u] ^= 1;
      mbar_wait(&bs[2], full_phase[u]);   // and its magnitudes are in shared memory
      full_phase[u] ^= 1;
      // e_i = mask_i * |Y| against every source magnitude; lane j owns bins j + 32 r, two of them per packed operation
#pragma unroll
      for (int r = 0; r < ((exp & 1) ? 1 : 8); ++r) {
        const float2 ov = make_float2(yrow[64 * r], yrow[64 * r + 32]);
        float2 x[K];
#pragma unroll
        for (int j = 0; j < K; ++j) x[j] = make_float2(mg[j * kWsMagRow + 64 * r], mg[j * kWsMagRow + 64 * r + 32]);
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float2 e = rf::mul2(make_float2(mrow[i * F + 64 * r], mrow[i * F + 64 * r + 32]), ov);
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const float2 d = rf::sub2(e, x[j]);
            acc[u][i * K + j] = rf::fma2(d, d, acc[u][i * K + j]);
          }
        }
      }
      {   // bin 512: one lane's worth (every lane reads the same words; lane 0's square counts)
        const float oy = yrow[rf::kHalf - lane];
        float x[K];
#pragma unroll
        for (int j = 0; j < K; ++j) x[j] = mg[j * kWsMagRow + rf::kHalf - lane];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float e = mrow[i * F + rf::kHalf - lane] * oy;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const float d = e - x[j];
            acc[u][i * K + j].x = fmaf(lane == 0 ? d : 0.f, d, acc[u][i * K + j].x);
          }
        }
      }
      mbar_arrive(&bs[3]);   // every lane after its own loads: the magnitude buffer is free again
      __syncwarp();          // every lane has read its rows: the landing areas may be overwritten
    }
    start_rows(u, q[u] + 1, bn, mn);
    b[u] = bn; m[u] = mn;
    ++q[u];
  };

  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int u = 0; u < U; ++u) start_rows(u, q[u], b[u], m[u]);
  for (;;) {
    bool any = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (q[u] < p_end[u]) {
        step(u);
        any = true;
      }
    }
    if (!any) break;
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (b_cur[u] >= 0) flush(u, b_cur[u]);
}

template <int K, bool PAIR>
int launch_fused_ws(const b2s_stft_plan* plan, const float* yabs, const float* sources, const float* mask,
                    const int64_t* meta, int64_t batch, int64_t samples, int64_t frames, int64_t pad_left,
                    float* loss, int32_t* perm, double* sse, void* workspace, cudaStream_t stream) {
  B2S_REQUIRE(plan->shift % 4 == 0, "the fused STFT->PIT kernel needs a shift that is a multiple of 4 (got %d)",
              plan->shift);
  B2S_REQUIRE(frames >= 1, "the fused STFT->PIT kernel needs at least one frame");
  B2S_REQUIRE(samples < ((int64_t)1 << 30) && frames < ((int64_t)1 << 20) && pad_left < ((int64_t)1 << 30),
              "signal too long for the fused STFT->PIT kernel (%lld samples)", (long long)samples);
  const int64_t total = batch * frames;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, kWsStreams), kNumSMs));
  const int64_t nstreams = std::min<int64_t>((int64_t)grid * kWsStreams, total);
  const int slots = (int)(ceil_div(frames * nstreams, total) + 2);
  constexpr size_t smem = sizeof(float) * kWsStreams * ws_stream_floats(K);
  static_assert(smem + 1024 <= 227 * 1024, "stream areas exceed the shared memory of an SM");
  auto kernel = stft_pit_ws_kernel<K, PAIR>;
  static bool configured[64] = {};
  if (!configured[plan->device & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[plan->device & 63] = true;
  }
  static const bool use_pdl = [] { const char* e = getenv("B2S_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kWsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  const int shift = plan->shift;
  const float4* table = plan->lane_fwd;
  const float* window = plan->awin;
  const float2* tab = plan->pair_tw;
  const char* ee = getenv("B2S_FUSED_WS_EXP");   // tuning builds only (results are wrong with any bit set)
  const int exp = ee ? atoi(ee) : 0;
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  B2S_CUDA(cudaLaunchKernelEx(&cfg, kernel, yabs, sources, mask, meta, batch, samples, frames, shift, pad_left, table,
                              slots, partial, counters, loss, perm, sse, window, tab, exp));
  B2S_LAUNCH_CHECK("stft_pit_ws_kernel");
  return B2S_OK;
}
