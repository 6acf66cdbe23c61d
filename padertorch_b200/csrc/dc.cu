// Deep-clustering affinity loss: one streaming pass builds the (E+K) x (E+K) Gram matrix of the
// stacked [embedding | target mask] channels per example, straight from the model's native 't e f'
// layout; the loss is a function of that matrix only.
// Reference: deep_clustering_loss, padertorch/ops/losses/source_separation.py:13-31 (three tall-skinny
// einsums) called per example after two transpose copies in padertorch/contrib/tcl/dc.py:76-84.
//
// Register blocking: a warp owns one 8 x 8 block pair of the Gram matrix (64 accumulators per lane),
// lanes own consecutive time-frequency points (coalesced along f); the warps of a CTA walk the same
// point tiles, so the repeated channel reads hit L1.  fp32 FFMA on CUDA cores: 5.5 FLOP/B keeps this
// HBM-bound, tensor cores would need a 3x split to hold fp32 accuracy and lose to the padding.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "rfft_packed.cuh"   // packed fp32x2 helpers (rf::fma2, rf::bcast)

using namespace b2s;

namespace {

constexpr int BS = 8;            // Gram block edge
constexpr int kMaxWarps = 8;
constexpr int kDcBinBlock = 2048;   // bins per work unit (long [N, E] inputs are split along the points)

struct DcGrid { int blocks, pairs, warps, nchunks, cp; bool tiled; };

DcGrid dc_grid(int64_t batch, int64_t max_frames, int64_t bins, int channels, bool unit_bin_stride = true) {
  DcGrid g;
  g.tiled = channels <= 24 && unit_bin_stride;
  if (g.tiled) {
    g.blocks = 3; g.cp = 24; g.pairs = 6; g.warps = 6;
    const int64_t points = std::max<int64_t>(1, max_frames * bins);
    int64_t c = (int64_t)kNumSMs * 2 / std::max<int64_t>(1, batch);
    c = std::min<int64_t>(c, std::max<int64_t>(1, points / 512));
    g.nchunks = (int)std::max<int64_t>(1, c);
    return g;
  }
  g.blocks = (channels + BS - 1) / BS;
  g.cp = g.blocks * BS;
  g.pairs = g.blocks * (g.blocks + 1) / 2;
  g.warps = std::min(g.pairs, kMaxWarps);
  const int per_sm = std::max(1, std::min(2048 / (32 * g.warps), 65536 / (144 * 32 * g.warps)));
  const int64_t capacity = (int64_t)kNumSMs * per_sm;
  const int64_t units = std::max<int64_t>(1, max_frames * ceil_div(std::max<int64_t>(bins, 1), (int64_t)kDcBinBlock));
  int64_t c = capacity / std::max<int64_t>(1, batch);
  c = std::min<int64_t>(c, units);
  g.nchunks = (int)std::max<int64_t>(1, c);
  return g;
}

struct Strides { int64_t t, c, f; };

// Ticket + fixed-order fold of the chunk partials -> gram[b][C][C] and the loss (all threads call it).
// Sums of 32 per-lane values across the warp, value l delivered to lane l: five halving steps in which a lane keeps
// the half of its values its own lane bit selects, sends the other half to its partner and adds what it receives --
// 31 shuffles for 32 sums instead of 160, and afterwards every lane holds one result (parallel stores).
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void dc_finish(int b, int nchunks, int cp, int C, int E, int64_t N,
                                          double* __restrict__ partial, int* __restrict__ counters,
                                          double* __restrict__ gram, float* __restrict__ loss, int chunk_slots = 0,
                                          float* __restrict__ mean = nullptr, int* __restrict__ done = nullptr,
                                          int batch = 0) {
  if (chunk_slots == 0) chunk_slots = nchunks;   // partial matrices reserved per example
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(counters + b, 1) == nchunks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: fold chunks in order -> gram[b][C][C]; loss = (|A|^2 - 2|Cx|^2 + |D|^2) / N^2
  __shared__ double red[32 * kMaxWarps];
  double local = 0.0;
  for (int idx = threadIdx.x; idx < C * C; idx += blockDim.x) {
    const int r = idx / C, c = idx - r * C;
    const double s = ordered_sum(partial + (int64_t)b * chunk_slots * cp * cp + r * cp + c, nchunks, (int64_t)cp * cp);
    gram[(int64_t)b * C * C + idx] = s;
    const bool re = r < E, ce = c < E;
    const double w = (re == ce) ? 1.0 : -1.0;  // the two mixed blocks together give -2 |V^T Y|^2
    local += w * s * s;
  }
  // fixed tree: lanes by xor-shuffle, then the warps' sums in warp order
  const double wsum = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) s += red[w];
    loss[b] = (float)(s / ((double)N * (double)N));
    counters[b] = 0;
  }
  // b2s_dc_forward_mean: the CTA that finishes the LAST example folds the batch mean (dc_loss of
  // DeepClusteringModel.review, tcl/dc.py:83-84) -- no separate reduction launch
  if (mean == nullptr) return;   // kernel-uniform
  __shared__ int s_all;
  if (threadIdx.x == 0) {
    __threadfence();
    s_all = atomicAdd(done, 1) == batch - 1;
  }
  __syncthreads();
  if (!s_all) return;
  __threadfence();
  if (threadIdx.x < 32) {   // fixed order: lane-strided partial sums, warp tree
    const volatile float* values = loss;
    double local = 0.0;
    for (int i = threadIdx.x; i < batch; i += 32) local += (double)values[i];
    local = warp_sum(local);
    if (threadIdx.x == 0) { *mean = (float)(local / (double)batch); *done = 0; }
  }
}


__global__ void __launch_bounds__(32 * kMaxWarps)
dc_gram_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
               const int64_t* __restrict__ meta, Strides se, Strides st, int nchunks, int64_t F, int E,
               int K, int blocks, int pairs, double* __restrict__ partial, int* __restrict__ counters,
               double* __restrict__ gram, float* __restrict__ loss) {
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = E + K, cp = blocks * BS;
  const int64_t T = meta[b * B2S_DC_META + 0];
  const float* e_ = emb + meta[b * B2S_DC_META + 1];
  const float* t_ = tgt + meta[b * B2S_DC_META + 2];
  const int64_t N = T * F;
  // work units = (frame, block of kDcBinBlock bins); this CTA owns the contiguous range [u0, u1)
  const int64_t fblocks = ceil_div(F, (int64_t)kDcBinBlock);
  const int64_t units = T * fblocks;
  const int64_t u0 = units * chunk / nchunks, u1 = units * (chunk + 1) / nchunks;
  double* mine = partial + ((int64_t)b * nchunks + chunk) * cp * cp;

  for (int pair = warp; pair < pairs; pair += nwarps) {
    // pair index -> (ba <= bb), row-major over the upper triangle
    int ba = 0, rem = pair;
    while (rem >= blocks - ba) { rem -= blocks - ba; ++ba; }
    const int bb = ba + rem;
    // per-channel base pointers (nullptr: padding channel) and the per-channel point strides
    const float* pa[BS]; const float* pb[BS];
    bool ea[BS], eb[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) {
      const int ca = ba * BS + i, cb = bb * BS + i;
      ea[i] = ca < E; eb[i] = cb < E;
      pa[i] = ca < E ? e_ + ca * se.c : (ca < C ? t_ + (ca - E) * st.c : nullptr);
      pb[i] = cb < E ? e_ + cb * se.c : (cb < C ? t_ + (cb - E) * st.c : nullptr);
    }
    float acc[BS][BS];
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) acc[i][j] = 0.f;
    for (int64_t u = u0; u < u1; ++u) {
      const int64_t t = u / fblocks;
      const int64_t fbeg = (u - t * fblocks) * kDcBinBlock, fend = min(F, fbeg + kDcBinBlock);
      const int64_t oe0 = t * se.t, ot0 = t * st.t;
      for (int64_t f = fbeg + lane; f < fend; f += 32) {
        const int64_t oe = oe0 + f * se.f, ot = ot0 + f * st.f;
        float va[BS], vb[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) va[i] = pa[i] ? __ldg(pa[i] + (ea[i] ? oe : ot)) : 0.f;
        if (ba == bb) {
#pragma unroll
          for (int i = 0; i < BS; ++i) vb[i] = va[i];
        } else {
#pragma unroll
          for (int i = 0; i < BS; ++i) vb[i] = pb[i] ? __ldg(pb[i] + (eb[i] ? oe : ot)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
          for (int j = 0; j < BS; ++j) acc[i][j] = fmaf(va[i], vb[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const float s = warp_sum(acc[i][j]);
        if (lane == 0) {
          mine[(ba * BS + i) * cp + bb * BS + j] = (double)s;
          mine[(bb * BS + j) * cp + ba * BS + i] = (double)s;
        }
      }
  }
  dc_finish(b, nchunks, cp, C, E, N, partial, counters, gram, loss);
}

// ------------------------------------------------------------------------------------------- tiled Gram
// Fast path for E + K <= 24 channels with unit stride along the bins (the model's 't e f' layout):
// tiles of 128 consecutive time-frequency points x 24 channels are copied into shared memory in their
// native [channel][point] orientation (4-byte zero-filling cp.async, double buffered), and each of the
// 6 warps owns one 8 x 8 block pair: a lane reads 4 consecutive points of a channel with ONE LDS.128, so
// 16 shared loads feed 256 FMAs.
constexpr int kTP = 128;             // points per tile
constexpr int kTC = 24;              // channel rows per tile (3 blocks of 8)
constexpr int kTiledWarps = 6;       // block pairs of a 3 x 3 upper triangle (compute warps)
constexpr int kTiledThreads = 256;   // 8 warps load, the first 6 also compute

__device__ __forceinline__ void cp_async_4_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kTiledThreads, 2)
dc_gram_tiled_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
                     const int64_t* __restrict__ meta, Strides se, Strides st, int nchunks, int64_t F, int E,
                     int K, double* __restrict__ partial, int* __restrict__ counters,
                     double* __restrict__ gram, float* __restrict__ loss) {
  __shared__ __align__(16) float tile[2][kTC][kTP];
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = E + K;
  const int64_t T = meta[b * B2S_DC_META + 0];
  const float* e_ = emb + meta[b * B2S_DC_META + 1];
  const float* t_ = tgt + meta[b * B2S_DC_META + 2];
  const int64_t N = T * F;
  const int64_t q0 = N * chunk / nchunks, q1 = N * (chunk + 1) / nchunks;

  // loader: 256 threads = 128 point columns x 2 channel phases; thread (p, h) copies channels h, h+2, ...
  // of its point: one division per tile, then a pointer increment per element
  const int p = threadIdx.x & (kTP - 1), h = threadIdx.x >> 7;
  auto issue_tile = [&](int64_t qt, int buf) {
    const int64_t q = qt + p;
    const bool ok = q < q1;
    const int64_t t = q / F, f = q - t * F;
    const float* pe = e_ + t * se.t + f + h * se.c;
    const float* pt = t_ + t * st.t + f;
    float* dst = &tile[buf][h][p];
    int c = h;
    for (; c < E; c += 2, pe += 2 * se.c, dst += 2 * kTP) cp_async_4_zfill(dst, ok ? pe : e_, ok);
    pt += (c - E) * st.c;
    for (; c < C; c += 2, pt += 2 * st.c, dst += 2 * kTP) cp_async_4_zfill(dst, ok ? pt : e_, ok);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // padding channel rows stay zero for the whole kernel
  for (int e = threadIdx.x; e < 2 * kTC * kTP; e += kTiledThreads) {
    const int c = (e >> 7) % kTC;
    if (c >= C) (&tile[0][0][0])[e] = 0.f;
  }
  // this warp's block pair (ba <= bb) of the 3 x 3 upper triangle: 00 01 02 11 12 22
  const int ba = warp < 3 ? 0 : (warp < 5 ? 1 : 2);
  const int bb = warp < 3 ? warp : (warp < 5 ? warp - 2 : 2);
  const bool active = warp < kTiledWarps && ba * BS < C && bb * BS < C;
  float acc[BS][BS];
#pragma unroll
  for (int i = 0; i < BS; ++i)
#pragma unroll
    for (int j = 0; j < BS; ++j) acc[i][j] = 0.f;

  if (q0 < q1) issue_tile(q0, 0);
  int cur = 0;
  for (int64_t qt = q0; qt < q1; qt += kTP, cur ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                               // tile `cur` complete and visible; other buffer free
    if (qt + kTP < q1) issue_tile(qt + kTP, cur ^ 1);
    if (active) {
      float4 va[BS];
#pragma unroll
      for (int i = 0; i < BS; ++i) va[i] = *reinterpret_cast<const float4*>(&tile[cur][ba * BS + i][4 * lane]);
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        // one row of the B block at a time keeps the live set at 64 accumulators + 36 operands
        const float4 vb = ba == bb ? va[j] : *reinterpret_cast<const float4*>(&tile[cur][bb * BS + j][4 * lane]);
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          float a = acc[i][j];
          a = fmaf(va[i].x, vb.x, a);
          a = fmaf(va[i].y, vb.y, a);
          a = fmaf(va[i].z, vb.z, a);
          a = fmaf(va[i].w, vb.w, a);
          acc[i][j] = a;
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  double* mine = partial + ((int64_t)b * nchunks + chunk) * kTC * kTC;
  if (warp < kTiledWarps) {
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const float s = warp_sum(acc[i][j]);
        if (lane == 0) {
          mine[(ba * BS + i) * kTC + bb * BS + j] = (double)s;
          mine[(bb * BS + j) * kTC + ba * BS + i] = (double)s;
        }
      }
  }
  dc_finish(b, nchunks, kTC, C, E, N, partial, counters, gram, loss);
}

// ------------------------------------------------------------------------------------------- frame-tiled Gram
// Fast path for the model's contiguous 't e f' layout (channel stride = bins, bin stride 1) with E + K <= 24:
// the [E][F] embedding block and the [K][F] target block of one frame are each ONE contiguous span of global
// memory, fetched by one TMA bulk copy each (cp.async.bulk -> mbarrier; the enclosing 16-byte aligned range
// lands in shared memory, the block sits at the source's misalignment), double buffered per CTA.  Each of the 6
// warps owns one 8 x 8 block pair; lanes own consecutive bins, so every operand is a conflict-free LDS.32
// whatever the rows' alignment.  Replaces 2816 four-byte cp.async per 128 points (about 8 LSU cycles per warp
// instruction: the copy, not the arithmetic, bounded the tiled kernel) by two descriptor-less bulk copies.
constexpr int kFrWarps = 6;
__host__ __device__ inline int frame_area(int rows, int F) { return (rows * F + 3 + 3) / 4 * 4 + 32; }

// One frame of block pair (BA, BB): lanes own bins lane, lane + 32, ...; with compile-time geometry every
// operand address is base register + immediate.
// The accumulators are packed pairs (acc[i][jj] = entries (i, 2 jj) and (i, 2 jj + 1)): one FFMA2 with a broadcast
// operand per pair instead of two FFMA -- the kernel is bound by issue slots (6 500 instructions per frame and CTA
// with scalar FFMA), not by HBM or the FMA pipe; a diagonal block computes the pairs that touch its upper triangle
// (20 instead of 36 scalar products' worth of instructions).
template <int FT, int ET, int KT, int BA, int BB>
__device__ __forceinline__ void gram_frame_block(const float* be_, const float* bt_, const float* zrow, int F_rt,
                                                 int E_rt, int K_rt, int lane, float2 (&acc)[BS][BS / 2]) {
  const int F = FT ? FT : F_rt, E = ET ? ET : E_rt, C = E + (KT ? KT : K_rt);
  constexpr bool diag = BA == BB;
  auto row = [&](int ch, int off) -> const float* {
    return (ch < E ? be_ + ch * F : (ch < C ? bt_ + (ch - E) * F : zrow)) + off;
  };
  auto step = [&](int off) {
    float va[BS];
    float2 vb[BS / 2];
#pragma unroll
    for (int i = 0; i < BS; ++i) va[i] = row(BA * BS + i, off)[0];
#pragma unroll
    for (int jj = 0; jj < BS / 2; ++jj)
      vb[jj] = diag ? make_float2(va[2 * jj], va[2 * jj + 1])
                    : make_float2(row(BB * BS + 2 * jj, off)[0], row(BB * BS + 2 * jj + 1, off)[0]);
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int jj = diag ? i / 2 : 0; jj < BS / 2; ++jj) acc[i][jj] = rf::fma2(rf::bcast(va[i]), vb[jj], acc[i][jj]);
  };
  const int full_steps = F / 32;
#pragma unroll 4
  for (int j = 0; j < full_steps; ++j) step(32 * j);
  if (lane + 32 * full_steps < F) step(32 * full_steps);   // the F % 32 last bins
}

// FT / ET / KT != 0: bins / embedding channels / sources known at compile time (513 / 20 / 2)
template <int FT, int ET, int KT>
__global__ void __launch_bounds__(32 * kFrWarps, 2)
dc_gram_frame_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
                     const int64_t* __restrict__ meta, int64_t se_t, int64_t st_t, int nchunks, int F_rt, int E_rt,
                     int K_rt, double* __restrict__ partial, int* __restrict__ counters,
                     double* __restrict__ gram, float* __restrict__ loss, float* __restrict__ mean) {
  extern __shared__ __align__(16) float fsm[];   // [2][area_e + area_t] frame buffers, then a row of zeros
  __shared__ __align__(8) uint64_t full[2];
  const int F = FT ? FT : F_rt, E = ET ? ET : E_rt, K = KT ? KT : K_rt;
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = E + K;
  const int area_e = frame_area(E, F), area_t = frame_area(K, F), buf_floats = area_e + area_t;
  float* zrow = fsm + 2 * buf_floats;
  const int64_t T = meta[b * B2S_DC_META + 0];
  const float* e_ = emb + meta[b * B2S_DC_META + 1];
  const float* t_ = tgt + meta[b * B2S_DC_META + 2];
  const int64_t N = T * F;
  const int t0 = (int)(T * chunk / nchunks), t1 = (int)(T * (chunk + 1) / nchunks);
  for (int i = threadIdx.x; i < (F + 31) / 32 * 32; i += blockDim.x) zrow[i] = 0.f;
  if (threadIdx.x == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {   // thread 0
    const uintptr_t ae = reinterpret_cast<uintptr_t>(e_ + (int64_t)t * se_t);
    const uintptr_t at = reinterpret_cast<uintptr_t>(t_ + (int64_t)t * st_t);
    const unsigned be = (unsigned)(((ae & 15) + (size_t)E * F * 4 + 15) & ~(size_t)15);
    const unsigned bt = (unsigned)(((at & 15) + (size_t)K * F * 4 + 15) & ~(size_t)15);
    tma::fence_proxy_async();   // the buffer was last read through the generic proxy
    tma::mbar_expect_tx(&full[s], be + bt);
    tma::bulk_g2s(fsm + s * buf_floats, reinterpret_cast<const void*>(ae & ~(uintptr_t)15), be, &full[s]);
    tma::bulk_g2s(fsm + s * buf_floats + area_e, reinterpret_cast<const void*>(at & ~(uintptr_t)15), bt, &full[s]);
  };
  if (threadIdx.x == 0) {
    if (t0 < t1) issue(t0, 0);
    if (t0 + 1 < t1) issue(t0 + 1, 1);
  }
  // this warp's block pair (ba <= bb) of the 3 x 3 upper triangle: 00 01 02 11 12 22
  const int ba = warp < 3 ? 0 : (warp < 5 ? 1 : 2);
  const int bb = warp < 3 ? warp : (warp < 5 ? warp - 2 : 2);
  const bool active = ba * BS < C && bb * BS < C;
  const bool diag = ba == bb;
  float2 acc[BS][BS / 2];   // packed pairs of Gram entries (i, 2 jj), (i, 2 jj + 1)
#pragma unroll
  for (int i = 0; i < BS; ++i)
#pragma unroll
    for (int j = 0; j < BS / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  for (int t = t0; t < t1; ++t) {
    const int s = (t - t0) & 1;
    tma::mbar_wait(&full[s], (unsigned)((t - t0) >> 1) & 1u);
    if (active) {
      const float* be_ = fsm + s * buf_floats + (int)((reinterpret_cast<uintptr_t>(e_ + (int64_t)t * se_t) & 15) >> 2) + lane;
      const float* bt_ = fsm + s * buf_floats + area_e + (int)((reinterpret_cast<uintptr_t>(t_ + (int64_t)t * st_t) & 15) >> 2) + lane;
      const float* z_ = zrow + lane;
      switch (warp) {   // warp-uniform
        case 0: gram_frame_block<FT, ET, KT, 0, 0>(be_, bt_, z_, F, E, K, lane, acc); break;
        case 1: gram_frame_block<FT, ET, KT, 0, 1>(be_, bt_, z_, F, E, K, lane, acc); break;
        case 2: gram_frame_block<FT, ET, KT, 0, 2>(be_, bt_, z_, F, E, K, lane, acc); break;
        case 3: gram_frame_block<FT, ET, KT, 1, 1>(be_, bt_, z_, F, E, K, lane, acc); break;
        case 4: gram_frame_block<FT, ET, KT, 1, 2>(be_, bt_, z_, F, E, K, lane, acc); break;
        default: gram_frame_block<FT, ET, KT, 2, 2>(be_, bt_, z_, F, E, K, lane, acc); break;
      }
    }
    __syncthreads();   // every warp is done with buffer s
    if (threadIdx.x == 0 && t + 2 < t1) issue(t + 2, s);
  }
  double* mine = partial + ((int64_t)b * nchunks + chunk) * kTC * kTC;
  // the block's 64 sums in two transposed reductions: lane l receives entries l and 32 + l (row-major i * 8 + j) and
  // stores them (and their mirror images) itself
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int i = (32 * half + e) / BS, j = (32 * half + e) % BS;
      v[e] = (j & 1) ? acc[i][j / 2].y : acc[i][j / 2].x;
    }
    const float sum = warp_transpose_sum32(v, lane);
    const int i = (32 * half + lane) / BS, j = (32 * half + lane) % BS;
    if (!(diag && j < i)) {   // the lower triangle of a diagonal block is its mirror image
      mine[(ba * BS + i) * kTC + bb * BS + j] = (double)sum;
      mine[(bb * BS + j) * kTC + ba * BS + i] = (double)sum;
    }
  }
  dc_finish(b, nchunks, kTC, C, E, N, partial, counters, gram, loss, 0, mean, counters + (kMaxTickets - 1), (int)gridDim.x);
}

// ------------------------------------------------------------------------------------------- ring form of the frame kernel
// dc_gram_frame_kernel pays for two things its arithmetic does not need: (1) its six warps carry unequal work (a
// diagonal block is 20 packed products per step, an off-diagonal one 32) and meet at a block barrier after EVERY
// frame, and (2) each block pair loads its own 16 operand rows (96 LDS per 32 bins and CTA).  Here a warp owns one
// diagonal block AND the off-diagonal block that shares its row block -- (0,0)+(0,1), (1,1)+(1,2), (2,2)+(2,0): 52
// packed products per step for every warp, 16 operand loads for both blocks (48 LDS per 32 bins) -- and four such
// role triples (12 warps, one CTA per SM) share the 32-bin steps of the frames round robin over a running step
// counter, so all warps carry the same work whatever the number of bins.  Frames travel through a ring of four
// whole-frame stages (TMA bulk copies as above, `full` mbarrier per stage); a warp releases a stage by one non-blocking
// arrival on the stage's named barrier (bar.arrive) and runs on; no block-wide wait exists in the frame loop.  Warp 0
// refills the stage of frame f - 1 with frame f + 3 when it starts frame f (bar.sync on that stage's barrier: three
// frames of slack for everyone else).
constexpr int kRgGroups = 4, kRgWarps = 3 * kRgGroups, kRgStages = 4;
// Length-balanced chunking of a ragged batch on a one-dimensional grid of P slots: the smallest chunk length L with
// sum_b ceil(T_b / L) <= P is found by bisection between sum(T) / P and sum(T) / (P - B), example b is split into
// n_b = ceil(T_b / L) equal chunks -- no CTA carries more than L frames whatever the spread of the lengths (a uniform
// number of chunks per example makes the CTAs of the longest example 1.6 x the mean at lengths uniform in 2 .. 8 s),
// and all chunks are resident in ONE wave.  The batch's meta rows are staged in shared memory on the way (the slot's
// own row is read from there: one global round trip in all).  Every thread calls it; slots beyond sum(n_b) return
// b = -1.  Needs B <= P, B <= kBalMaxBatch.
constexpr int kBalMaxBatch = 160;
struct ChunkSlot { int b, chunk, nchunks; int64_t row[B2S_DC_META]; };   // row: the example's meta row
__device__ inline ChunkSlot balanced_chunk_slot(const int64_t* __restrict__ meta, int B, int P, int max_chunks, int slot) {
  __shared__ int64_t s_meta[kBalMaxBatch * B2S_DC_META];
  __shared__ ChunkSlot s_slot;
  for (int i = threadIdx.x; i < B * B2S_DC_META; i += blockDim.x) s_meta[i] = meta[i];
  if (threadIdx.x == 0) s_slot.b = -1;
  bool same = true;   // the lengths this thread looks at equal example 0's (read from global memory again: the
                      // staged copies are not visible before the barrier; same addresses, so the loads merge)
  for (int b = threadIdx.x; b < B; b += blockDim.x) same = same && meta[(int64_t)b * B2S_DC_META] == meta[0];
  if (__syncthreads_and(same)) {
    // equal lengths (dense batches): closed form, no reductions -- L = ceil(T / floor(P / B))
    const int T = (int)s_meta[0];
    const int m = max(1, min(max_chunks, P / B));
    const int L = max(8, (T + m - 1) / m);
    const int n = max(1, min(max_chunks, (T + L - 1) / L));
    // example index fastest, as the (batch, chunks) grid enumerates its CTAs (consecutive slots on consecutive
    // chunks of ONE example measured 1 us slower at batch 16 x 253 frames)
    ChunkSlot r;
    r.chunk = slot / B;
    r.b = slot - r.chunk * B;
    r.nchunks = n;
    if (r.chunk >= n) { r.b = -1; return r; }
#pragma unroll
    for (int i = 0; i < B2S_DC_META; ++i) r.row[i] = s_meta[r.b * B2S_DC_META + i];
    return r;
  }
  if (threadIdx.x < 32) {   // warp 0
    const int lane = threadIdx.x;
    int tmax = 0;
    long long tsum = 0;
    for (int b = lane; b < B; b += 32) { const int T = (int)s_meta[b * B2S_DC_META]; tmax = max(tmax, T); tsum += T; }
    for (int off = 16; off >= 1; off >>= 1) {
      tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
      tsum += __shfl_xor_sync(0xffffffffu, tsum, off);
    }
    auto chunks_of = [&](int T, int L) { return max(1, min(max_chunks, (T + L - 1) / L)); };
    // at least eight frames per chunk; sum ceil(T_b / L) >= sum(T) / L and <= sum(T) / L + B bound L from both sides
    int lo = (int)max((long long)8, (tsum + P - 1) / P);
    int hi = max(lo, P > B ? (int)min((long long)tmax, (tsum + P - B - 1) / (P - B)) : tmax);
    while (lo < hi) {                // warp-uniform
      const int mid = (lo + hi) >> 1;
      int cnt = 0;
      for (int b = lane; b < B; b += 32) cnt += chunks_of((int)s_meta[b * B2S_DC_META], mid);
      for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
      if (cnt <= P) hi = mid; else lo = mid + 1;
    }
    int base = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {   // running sum over the examples, 32 at a time
      const int b = b0 + lane;
      const int n = b < B ? chunks_of((int)s_meta[b * B2S_DC_META], lo) : 0;
      int incl = n;
      for (int off = 1; off < 32; off <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += up;
      }
      const int first = base + incl - n;
      if (b < B && slot >= first && slot < first + n) {
        s_slot.b = b; s_slot.chunk = slot - first; s_slot.nchunks = n;
#pragma unroll
        for (int i = 0; i < B2S_DC_META; ++i) s_slot.row[i] = s_meta[b * B2S_DC_META + i];
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
      if (base > slot) break;   // warp-uniform
    }
  }
  __syncthreads();
  return s_slot;
}

// One 32-bin step of the blocks (BA, BA) [upper triangle] and (BA, BB).
template <int FT, int ET, int KT, int BA, int BB>
__device__ __forceinline__ void gram_ring_step(const float* be_, const float* bt_, const float* zrow, int F, int E,
                                               int C, int off, float2 (&accd)[BS][BS / 2], float2 (&acco)[BS][BS / 2]) {
  auto row = [&](int ch) -> const float* {
    return (ch < E ? be_ + ch * F : (ch < C ? bt_ + (ch - E) * F : zrow)) + off;
  };
  float va[BS];
  float2 vb[BS / 2];
#pragma unroll
  for (int i = 0; i < BS; ++i) va[i] = row(BA * BS + i)[0];
#pragma unroll
  for (int jj = 0; jj < BS / 2; ++jj) vb[jj] = make_float2(row(BB * BS + 2 * jj)[0], row(BB * BS + 2 * jj + 1)[0]);
#pragma unroll
  for (int i = 0; i < BS; ++i) {
#pragma unroll
    for (int jj = i / 2; jj < BS / 2; ++jj)
      accd[i][jj] = rf::fma2(rf::bcast(va[i]), make_float2(va[2 * jj], va[2 * jj + 1]), accd[i][jj]);
#pragma unroll
    for (int jj = 0; jj < BS / 2; ++jj) acco[i][jj] = rf::fma2(rf::bcast(va[i]), vb[jj], acco[i][jj]);
  }
}

template <int FT, int ET, int KT>
__global__ void __launch_bounds__(32 * kRgWarps, 1)
dc_gram_ring_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
                    const int64_t* __restrict__ meta, int64_t se_t, int64_t st_t, int batch, int balance_ctas,
                    int chunk_slots, int F_rt, int E_rt, int K_rt, double* __restrict__ partial,
                    int* __restrict__ counters, double* __restrict__ gram, float* __restrict__ loss,
                    float* __restrict__ mean) {
  extern __shared__ __align__(16) float fsm[];   // [kRgStages][area_e + area_t] frame stages, then a row of zeros
  __shared__ __align__(8) uint64_t full[kRgStages];
  const int F = FT ? FT : F_rt, E = ET ? ET : E_rt, K = KT ? KT : K_rt;
  // balance_ctas != 0: one-dimensional grid, chunks per example in proportion to its length; else grid (batch, chunks)
  {   // barriers and the row of zeros first: independent of the slot, visible after the barriers below
    const int area = frame_area(E, F) + frame_area(K, F);
    for (int i = threadIdx.x; i < (F + 31) / 32 * 32; i += blockDim.x) fsm[kRgStages * area + i] = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < kRgStages; ++s) {
        tma::mbar_init(&full[s], 1);
      }
      tma::fence_mbar_init();
    }
  }
  int b = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
  int64_t row[3];
  if (balance_ctas) {
    const ChunkSlot slot = balanced_chunk_slot(meta, batch, balance_ctas, chunk_slots, (int)blockIdx.x);
    if (slot.b < 0) return;
    b = slot.b; chunk = slot.chunk; nchunks = slot.nchunks;
#pragma unroll
    for (int i = 0; i < 3; ++i) row[i] = slot.row[i];
  } else {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 3; ++i) row[i] = meta[(int64_t)b * B2S_DC_META + i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = E + K;
  const int area_e = frame_area(E, F), area_t = frame_area(K, F), buf_floats = area_e + area_t;
  float* zrow = fsm + kRgStages * buf_floats;
  const int64_t T = row[0];
  const float* e_ = emb + row[1];
  const float* t_ = tgt + row[2];
  const int64_t N = T * F;
  const int t0 = (int)(T * chunk / nchunks), t1 = (int)(T * (chunk + 1) / nchunks);
  const int n = t1 - t0;
  auto issue = [&](int t, int s) {   // one lane
    const uintptr_t ae = reinterpret_cast<uintptr_t>(e_ + (int64_t)t * se_t);
    const uintptr_t at = reinterpret_cast<uintptr_t>(t_ + (int64_t)t * st_t);
    const unsigned be = (unsigned)(((ae & 15) + (size_t)E * F * 4 + 15) & ~(size_t)15);
    const unsigned bt = (unsigned)(((at & 15) + (size_t)K * F * 4 + 15) & ~(size_t)15);
    tma::fence_proxy_async();   // the stage was last read through the generic proxy
    tma::mbar_expect_tx(&full[s], be + bt);
    tma::bulk_g2s(fsm + s * buf_floats, reinterpret_cast<const void*>(ae & ~(uintptr_t)15), be, &full[s]);
    tma::bulk_g2s(fsm + s * buf_floats + area_e, reinterpret_cast<const void*>(at & ~(uintptr_t)15), bt, &full[s]);
  };
  if (threadIdx.x == 0) {
    for (int f = 0; f < kRgStages && f < n; ++f) issue(t0 + f, f);
  }
  const int role = warp % 3, group = warp / 3;
  float2 accd[BS][BS / 2], acco[BS][BS / 2];
#pragma unroll
  for (int i = 0; i < BS; ++i)
#pragma unroll
    for (int j = 0; j < BS / 2; ++j) { accd[i][j] = make_float2(0.f, 0.f); acco[i][j] = make_float2(0.f, 0.f); }

  const int full_steps = F / 32;
  const bool tail = (F & 31) != 0;
  const int nsteps = full_steps + (tail ? 1 : 0);
  int first = group;   // this warp's first step of the current frame: (f * nsteps + j) % kRgGroups == group
  for (int f = 0; f < n; ++f) {
    const int s = f & (kRgStages - 1);
    if (warp == 0 && f >= 1 && f + kRgStages - 1 < n) {
      // stage r is free once the other eleven warps have arrived on its named barrier (producer / consumer form of
      // bar.arrive / bar.sync: the consumers do not wait); warp 0's own loads of frame f - 1 precede this in program order
      const int r = (f - 1) & (kRgStages - 1);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + r), "n"(32 * kRgWarps) : "memory");
      if (lane == 0) issue(t0 + f + kRgStages - 1, r);
      __syncwarp();
    }
    tma::mbar_wait(&full[s], (unsigned)(f / kRgStages) & 1u);
    const int64_t t = t0 + f;
    const float* be_ = fsm + s * buf_floats + (int)((reinterpret_cast<uintptr_t>(e_ + t * se_t) & 15) >> 2) + lane;
    const float* bt_ = fsm + s * buf_floats + area_e + (int)((reinterpret_cast<uintptr_t>(t_ + t * st_t) & 15) >> 2) + lane;
    const float* z_ = zrow + lane;
    // this warp's steps of the frame: first, first + 4, ... (whole rounds unrolled: the next step's operand loads
    // overlap the current step's products), then the F % 32 last bins if that step is this warp's
    auto frame_steps = [&](auto stepf) {
      const int rounds = full_steps / kRgGroups;
#pragma unroll 4
      for (int q = 0; q < rounds; ++q) stepf(32 * (first + kRgGroups * q));
      const int j = first + kRgGroups * rounds;
      if (j < full_steps) stepf(32 * j);
      if (tail && (full_steps - first) % kRgGroups == 0 && lane + 32 * full_steps < F) stepf(32 * full_steps);
    };
    switch (role) {   // warp-uniform
      case 0: frame_steps([&](int off) { gram_ring_step<FT, ET, KT, 0, 1>(be_, bt_, z_, F, E, C, off, accd, acco); }); break;
      case 1: frame_steps([&](int off) { gram_ring_step<FT, ET, KT, 1, 2>(be_, bt_, z_, F, E, C, off, accd, acco); }); break;
      default: frame_steps([&](int off) { gram_ring_step<FT, ET, KT, 2, 0>(be_, bt_, z_, F, E, C, off, accd, acco); }); break;
    }
    first = (first + kRgGroups * nsteps - nsteps) % kRgGroups;   // (group - (f + 1) * nsteps) mod kRgGroups
    // release the stage if it will be refilled (frame f + kRgStages exists): one non-blocking arrival per warp
    if (warp != 0 && f + kRgStages < n) asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "n"(32 * kRgWarps) : "memory");
  }
  __syncthreads();   // every stage has been consumed by every warp: the stages are free for the reduction
  // Each warp reduces its 2 x 64 sums (transposed: lane l ends up with entries l and 32 + l of either block); the four
  // groups' sums meet in shared memory and are folded in group order.
  float* red = fsm;   // [kRgGroups][3 roles][4][32]
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int i = (32 * (q & 1) + e) / BS, jx = (32 * (q & 1) + e) % BS;
      const float2 a = q < 2 ? accd[i][jx / 2] : acco[i][jx / 2];
      v[e] = (jx & 1) ? a.y : a.x;
    }
    red[((group * 3 + role) * 4 + q) * 32 + lane] = warp_transpose_sum32(v, lane);
  }
  __syncthreads();
  double* mine = partial + ((int64_t)b * chunk_slots + chunk) * kTC * kTC;
  {
    const int r = warp % 3, q = (warp / 3);   // thread = (role r, quarter q, lane): 3 x 4 x 32 = blockDim
    double sum = 0.0;
#pragma unroll
    for (int g = 0; g < kRgGroups; ++g) sum += (double)red[((g * 3 + r) * 4 + q) * 32 + lane];
    const int ba = r, bb = q < 2 ? r : (r + 1) % 3;
    const int i = (32 * (q & 1) + lane) / BS, jx = (32 * (q & 1) + lane) % BS;
    if (!(q < 2 && jx < i)) {   // the lower triangle of a diagonal block is its mirror image
      mine[(ba * BS + i) * kTC + bb * BS + jx] = sum;
      mine[(bb * BS + jx) * kTC + ba * BS + i] = sum;
    }
  }
  dc_finish(b, nchunks, kTC, C, E, N, partial, counters, gram, loss, chunk_slots, mean, counters + (kMaxTickets - 1), batch);
}

// grad_V[p][e] = coef * ( sum_{c<E} V[p][c] G[c][e] - sum_{k} Y[p][k] G[e][E+k] ),  coef = 4 g / N^2.
// Thread = 2 points x one block of 8 output channels; the coefficient matrix sits in shared memory.
__global__ void __launch_bounds__(256)
dc_backward_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
                   const int64_t* __restrict__ meta, Strides se, Strides st, int nchunks, int64_t F, int E,
                   int K, const double* __restrict__ gram, const float* __restrict__ grad_loss, int64_t gl_stride, double gl_scale,
                   float* __restrict__ grad_emb) {
  extern __shared__ float coef[];  // [C][Ep], Ep = E rounded up to BS
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int C = E + K, eblocks = (E + BS - 1) / BS, Ep = eblocks * BS;
  const int64_t T = meta[b * B2S_DC_META + 0];
  const int64_t eoff = meta[b * B2S_DC_META + 1];
  const float* e_ = emb + eoff;
  const float* t_ = tgt + meta[b * B2S_DC_META + 2];
  float* g_ = grad_emb + meta[b * B2S_DC_META + 3];
  const int64_t N = T * F;
  const double scale = 4.0 * gl_scale * (double)grad_loss[(int64_t)b * gl_stride] / ((double)N * (double)N);
  for (int idx = threadIdx.x; idx < C * Ep; idx += blockDim.x) {
    const int c = idx / Ep, e = idx - c * Ep;
    double v = 0.0;
    if (e < E) v = (c < E ? 1.0 : -1.0) * scale * gram[(int64_t)b * C * C + c * C + e];
    coef[idx] = (float)v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // work units = (frame, block of 64 bins); this CTA owns [u0, u1); a warp item = (unit, output block)
  const int64_t fblocks = ceil_div(F, (int64_t)64);
  const int64_t units = T * fblocks;
  const int64_t u0 = units * chunk / nchunks, u1 = units * (chunk + 1) / nchunks;
  for (int64_t item = (u1 - u0) * 0 + warp; item < (u1 - u0) * eblocks; item += nwarps) {
    const int64_t u = u0 + item / eblocks;
    const int eb = (int)(item % eblocks);
    const int64_t t = u / fblocks;
    const int64_t fa = (u - t * fblocks) * 64 + lane, fb = fa + 32;
    const bool oka = fa < F, okb = fb < F;
    const float* ea_ = e_ + t * se.t + fa * se.f;
    const float* ta_ = t_ + t * st.t + fa * st.f;
    float ga[BS], gb[BS];
#pragma unroll
    for (int j = 0; j < BS; ++j) { ga[j] = 0.f; gb[j] = 0.f; }
    for (int c = 0; c < C; ++c) {
      const float* src = c < E ? ea_ + c * se.c : ta_ + (c - E) * st.c;
      const int64_t step = 32 * (c < E ? se.f : st.f);
      const float za = oka ? __ldg(src) : 0.f;
      const float zb = okb ? __ldg(src + step) : 0.f;
      const float4 c0 = *reinterpret_cast<const float4*>(coef + c * Ep + eb * BS);
      const float4 c1 = *reinterpret_cast<const float4*>(coef + c * Ep + eb * BS + 4);
      const float cf[BS] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int j = 0; j < BS; ++j) { ga[j] = fmaf(za, cf[j], ga[j]); gb[j] = fmaf(zb, cf[j], gb[j]); }
    }
    float* dst = g_ + t * se.t + fa * se.f;
#pragma unroll
    for (int j = 0; j < BS; ++j) {
      const int e = eb * BS + j;
      if (e < E) {
        if (oka) dst[e * se.c] = ga[j];
        if (okb) dst[e * se.c + 32 * se.f] = gb[j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- tiled backward
// grad[p][e] = sum_c z[p][c] * coef[c][e]: the same [channel][point] tiles as the forward (cp.async, double
// buffered); thread = 4 consecutive points x 5 output channels (E <= 20 -> 4 output groups x 32 point
// quads = 128 threads compute), operands by LDS.128 (tile) and broadcast LDS (coefficients).
constexpr int kBwdOut = 5;       // output channels per thread
constexpr int kBwdGroups = 4;    // output groups (covers E <= 20)

__global__ void __launch_bounds__(kTiledThreads, 2)
dc_backward_tiled_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
                         const int64_t* __restrict__ meta, Strides se, Strides st, int nchunks, int64_t F, int E,
                         int K, const double* __restrict__ gram, const float* __restrict__ grad_loss, int64_t gl_stride, double gl_scale,
                         float* __restrict__ grad_emb) {
  __shared__ __align__(16) float tile[2][kTC][kTP];
  __shared__ float coef[kTC][kBwdGroups * kBwdOut];   // [c][e], zero padded
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int C = E + K;
  const int64_t T = meta[b * B2S_DC_META + 0];
  const float* e_ = emb + meta[b * B2S_DC_META + 1];
  const float* t_ = tgt + meta[b * B2S_DC_META + 2];
  float* g_ = grad_emb + meta[b * B2S_DC_META + 3];
  const int64_t N = T * F;
  const int64_t q0 = N * chunk / nchunks, q1 = N * (chunk + 1) / nchunks;
  const double scale = 4.0 * gl_scale * (double)grad_loss[(int64_t)b * gl_stride] / ((double)N * (double)N);
  for (int idx = threadIdx.x; idx < kTC * kBwdGroups * kBwdOut; idx += kTiledThreads) {
    const int c = idx / (kBwdGroups * kBwdOut), e = idx - c * (kBwdGroups * kBwdOut);
    double v = 0.0;
    if (c < C && e < E) v = (c < E ? 1.0 : -1.0) * scale * gram[(int64_t)b * C * C + c * C + e];
    coef[c][e] = (float)v;
  }
  for (int e = threadIdx.x; e < 2 * kTC * kTP; e += kTiledThreads) {
    const int c = (e >> 7) % kTC;
    if (c >= C) (&tile[0][0][0])[e] = 0.f;
  }
  const int p = threadIdx.x & (kTP - 1), h = threadIdx.x >> 7;
  auto issue_tile = [&](int64_t qt, int buf) {
    const int64_t q = qt + p;
    const bool ok = q < q1;
    const int64_t t = q / F, f = q - t * F;
    const float* pe = e_ + t * se.t + f + h * se.c;
    const float* pt = t_ + t * st.t + f;
    float* dst = &tile[buf][h][p];
    int c = h;
    for (; c < E; c += 2, pe += 2 * se.c, dst += 2 * kTP) cp_async_4_zfill(dst, ok ? pe : e_, ok);
    pt += (c - E) * st.c;
    for (; c < C; c += 2, pt += 2 * st.c, dst += 2 * kTP) cp_async_4_zfill(dst, ok ? pt : e_, ok);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // compute threads: the first 128 -> (point quad = tid & 31, output group = tid >> 5)
  const int quad = threadIdx.x & 31, og = threadIdx.x >> 5;
  const bool active = og < kBwdGroups && og * kBwdOut < E;

  if (q0 < q1) issue_tile(q0, 0);
  int cur = 0;
  for (int64_t qt = q0; qt < q1; qt += kTP, cur ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (qt + kTP < q1) issue_tile(qt + kTP, cur ^ 1);
    if (active) {
      float4 g[kBwdOut];
#pragma unroll
      for (int j = 0; j < kBwdOut; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < C; ++c) {
        const float4 z = *reinterpret_cast<const float4*>(&tile[cur][c][4 * quad]);
#pragma unroll
        for (int j = 0; j < kBwdOut; ++j) {
          const float w = coef[c][og * kBwdOut + j];
          g[j].x = fmaf(z.x, w, g[j].x); g[j].y = fmaf(z.y, w, g[j].y);
          g[j].z = fmaf(z.z, w, g[j].z); g[j].w = fmaf(z.w, w, g[j].w);
        }
      }
      // 4 consecutive flattened points may straddle a frame boundary: store them one by one
      const float gv[kBwdOut][4] = {{g[0].x, g[0].y, g[0].z, g[0].w}, {g[1].x, g[1].y, g[1].z, g[1].w},
                                    {g[2].x, g[2].y, g[2].z, g[2].w}, {g[3].x, g[3].y, g[3].z, g[3].w},
                                    {g[4].x, g[4].y, g[4].z, g[4].w}};
      const int64_t qb = qt + 4 * quad;
      int64_t t = qb / F, f = qb - t * F;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (qb + i < q1) {
#pragma unroll
          for (int j = 0; j < kBwdOut; ++j) {
            const int e = og * kBwdOut + j;
            if (e < E) g_[t * se.t + e * se.c + f] = gv[j][i];
          }
        }
        if (++f == F) { f = 0; ++t; }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------- frame-tiled backward
// grad[p][e] = sum_c z[p][c] * coef[c][e] for the contiguous 't e f' layout: the frame's [E][F] / [K][F] blocks
// arrive by TMA bulk copies exactly as in dc_gram_frame_kernel (double buffered), the frame's [E][F] gradient
// block is assembled in shared memory at the destination's phase within 16 bytes and leaves as ONE TMA bulk
// store (double buffered; per-lane STG of rows that are only 4-byte aligned is several times slower).
// Warp = (output group of 5 channels, half of the frame's 64-bin steps): its 5 x C coefficients stay in
// registers, lanes own bins, two bins per lane in flight.  One CTA per SM.
constexpr int kFbOut = 5;

// FT / ET / KT != 0: bins / embedding channels / sources known at compile time (the reference's deep-clustering
// configuration 513 / 20 / 2): every shared-memory access of the inner loops becomes base register + immediate.
template <int FT, int ET, int KT>
__global__ void __launch_bounds__(ET ? 64 * ((ET + kFbOut - 1) / kFbOut) : 320, 1)
dc_backward_frame_kernel(const float* __restrict__ emb, const float* __restrict__ tgt,
                         const int64_t* __restrict__ meta, int64_t se_t, int64_t st_t, int nchunks_arg, int batch,
                         int balance_ctas, int F_rt, int E_rt,
                         int K_rt, const double* __restrict__ gram, const float* __restrict__ grad_loss, int64_t gl_stride, double gl_scale,
                         float* __restrict__ grad_emb) {
  const int F = FT ? FT : F_rt, E = ET ? ET : E_rt, K = KT ? KT : K_rt;
  extern __shared__ __align__(16) float fsm[];   // [2][area_e + area_t] inputs, [2][area_o] outputs
  __shared__ __align__(8) uint64_t full[2];
  __shared__ float coef[kTC][kTC + 1];           // [c][e], zero padded
  // balance_ctas != 0: one-dimensional grid, chunks per example in proportion to its length (balanced_chunk_slot)
  int b = blockIdx.x, chunk = blockIdx.y, nchunks = nchunks_arg;
  int64_t row[4];
  if (balance_ctas) {
    const ChunkSlot slot = balanced_chunk_slot(meta, batch, balance_ctas, 1 << 30, (int)blockIdx.x);
    if (slot.b < 0) return;
    b = slot.b; chunk = slot.chunk; nchunks = slot.nchunks;
#pragma unroll
    for (int i = 0; i < 4; ++i) row[i] = slot.row[i];
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) row[i] = meta[(int64_t)b * B2S_DC_META + i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = E + K;
  const int groups = (E + kFbOut - 1) / kFbOut;   // blockDim = 64 * groups
  const int og = warp % groups, half = warp / groups;
  const int area_e = frame_area(E, F), area_t = frame_area(K, F), buf_floats = area_e + area_t;
  const int area_o = frame_area(E, F);
  float* outs = fsm + 2 * buf_floats;
  const int64_t T = row[0];
  const float* e_ = emb + row[1];
  const float* t_ = tgt + row[2];
  float* g_ = grad_emb + row[3];
  const int64_t N = T * F;
  const int t0 = (int)(T * chunk / nchunks), t1 = (int)(T * (chunk + 1) / nchunks);
  const double scale = 4.0 * gl_scale * (double)grad_loss[(int64_t)b * gl_stride] / ((double)N * (double)N);
  for (int idx = threadIdx.x; idx < kTC * (kTC + 1); idx += blockDim.x) {
    const int c = idx / (kTC + 1), e = idx - c * (kTC + 1);
    double v = 0.0;
    if (c < C && e < E) v = (c < E ? 1.0 : -1.0) * scale * gram[(int64_t)b * C * C + c * C + e];
    coef[c][e] = (float)v;
  }
  if (threadIdx.x == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int t, int s) {   // thread 0
    const uintptr_t ae = reinterpret_cast<uintptr_t>(e_ + (int64_t)t * se_t);
    const uintptr_t at = reinterpret_cast<uintptr_t>(t_ + (int64_t)t * st_t);
    const unsigned be = (unsigned)(((ae & 15) + (size_t)E * F * 4 + 15) & ~(size_t)15);
    const unsigned bt = (unsigned)(((at & 15) + (size_t)K * F * 4 + 15) & ~(size_t)15);
    tma::fence_proxy_async();
    tma::mbar_expect_tx(&full[s], be + bt);
    tma::bulk_g2s(fsm + s * buf_floats, reinterpret_cast<const void*>(ae & ~(uintptr_t)15), be, &full[s]);
    tma::bulk_g2s(fsm + s * buf_floats + area_e, reinterpret_cast<const void*>(at & ~(uintptr_t)15), bt, &full[s]);
  };
  if (threadIdx.x == 0) {
    if (t0 < t1) issue(t0, 0);
    if (t0 + 1 < t1) issue(t0 + 1, 1);
  }
  // this warp's coefficients: cf[c][j] for output channel og * 5 + j (zero beyond E / C)
  float cf[kTC][kFbOut];
#pragma unroll
  for (int c = 0; c < kTC; ++c)
#pragma unroll
    for (int j = 0; j < kFbOut; ++j) cf[c][j] = coef[c][og * kFbOut + j];

  const int pairs = F / 64;            // steps of 64 bins (two per lane), alternating between the two halves
  const int rest0 = 64 * pairs;        // first bin of the remaining < 64 bins: one step of 32 per half
  const int n = E * F;
  for (int t = t0; t < t1; ++t) {
    const int s = (t - t0) & 1;
    float* gout = g_ + (int64_t)t * se_t;
    const int phase = (int)((reinterpret_cast<uintptr_t>(gout) & 15) >> 2);
    float* stage = outs + s * area_o + phase;   // stage[i] <-> gout[i]
    tma::mbar_wait(&full[s], (unsigned)((t - t0) >> 1) & 1u);
    const float* be_ = fsm + s * buf_floats + (int)((reinterpret_cast<uintptr_t>(e_ + (int64_t)t * se_t) & 15) >> 2) + lane;
    const float* bt_ = fsm + s * buf_floats + area_e + (int)((reinterpret_cast<uintptr_t>(t_ + (int64_t)t * st_t) & 15) >> 2) + lane;
    float* so = stage + og * kFbOut * F + lane;
    for (int ds = half; ds < pairs; ds += 2) {
      const float* ze = be_ + 64 * ds;
      const float* zt = bt_ + 64 * ds - E * F;
      auto row = [&](int c) { return c < E ? ze + c * F : zt + c * F; };
      float a0[kFbOut], a1[kFbOut];
#pragma unroll
      for (int j = 0; j < kFbOut; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
#pragma unroll
      for (int c = 0; c < kTC; ++c) {
        // compile-time geometry: the test folds away.  Run-time geometry: channels beyond C re-read channel 0 against
        // their zero coefficients instead of branching -- straight-line code whose operand loads the scheduler can
        // move ahead of the products (with a uniform branch per channel every load was consumed at once)
        if (ET != 0 && c >= C) continue;
        const float* r = row(ET != 0 || c < C ? c : 0);
        const float z0 = r[0], z1 = r[32];
#pragma unroll
        for (int j = 0; j < kFbOut; ++j) { a0[j] = fmaf(z0, cf[c][j], a0[j]); a1[j] = fmaf(z1, cf[c][j], a1[j]); }
      }
#pragma unroll
      for (int j = 0; j < kFbOut; ++j) {
        const int e = og * kFbOut + j;
        if (e < E) {
          so[j * F + 64 * ds] = a0[j];
          so[j * F + 64 * ds + 32] = a1[j];
        }
      }
    }
    {
      const int f = rest0 + 32 * half + lane;   // the F % 64 last bins
      if (f < F) {
        float a0[kFbOut];
#pragma unroll
        for (int j = 0; j < kFbOut; ++j) a0[j] = 0.f;
#pragma unroll
        for (int c = 0; c < kTC; ++c) {
          if (c < C) {
            const float z0 = (c < E ? be_ + c * F : bt_ + (c - E) * F)[rest0 + 32 * half];
#pragma unroll
            for (int j = 0; j < kFbOut; ++j) a0[j] = fmaf(z0, cf[c][j], a0[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < kFbOut; ++j) {
          const int e = og * kFbOut + j;
          if (e < E) stage[e * F + f] = a0[j];
        }
      }
    }
    tma::fence_proxy_async();                          // the block was written through the generic proxy
    if (threadIdx.x == 0) tma::bulk_wait_read<0>();    // earlier stores have read their staging buffers
    __syncthreads();                                   // inputs of buffer s consumed, block s complete
    const int head = (4 - phase) & 3, mid = (n - head) & ~3, tail = n - head - mid;
    if (threadIdx.x == 0) {
      if (t + 2 < t1) issue(t + 2, s);
      tma::bulk_s2g(gout + head, stage + head, (unsigned)mid * 4u);
      tma::bulk_commit();
    }
    if (warp == 1) {
      if (lane < head) gout[lane] = stage[lane];
      if (lane < tail) gout[head + mid + lane] = stage[head + mid + lane];
    }
  }
  if (threadIdx.x == 0) tma::bulk_wait<0>();   // shared memory must outlive the last store
}

}  // namespace

// Geometries (bins, embedding channels, sources) with compile-time instances of the ring Gram kernel and the frame
// backward kernel: the reference's deep-clustering model (tcl/dc.py:8-15: F = 257, E = 20) and the 1024-point STFT of
// BASELINE.json's configurations, two and three speakers.
#define B2S_DC_GEOMETRIES(X) X(513, 20, 2) X(257, 20, 2) X(513, 20, 3) X(257, 20, 3)

extern "C" {

int64_t b2s_dc_workspace_bytes(int64_t batch, int64_t max_frames, int64_t bins, int channels) {
  if (batch <= 0 || channels <= 0) return 16;
  const DcGrid g = dc_grid(batch, max_frames, bins, channels, true);
  const DcGrid h = dc_grid(batch, max_frames, bins, channels, false);
  const int64_t cells = std::max<int64_t>((int64_t)g.nchunks * g.cp * g.cp, (int64_t)h.nchunks * h.cp * h.cp);
  return kTicketBytes + (int64_t)sizeof(double) * batch * cells + 16;
}

}  // extern "C"

namespace {
// mean of loss[0 .. batch): the fallback of b2s_dc_forward_mean behind the kernels that do not fold it themselves
__global__ void __launch_bounds__(32)
dc_mean_kernel(const float* __restrict__ loss, int batch, float* __restrict__ mean) {
  double local = 0.0;
  for (int i = threadIdx.x; i < batch; i += 32) local += (double)loss[i];
  local = warp_sum(local);
  if (threadIdx.x == 0) *mean = (float)(local / (double)batch);
}

int dc_forward_impl(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                    int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                    const int64_t* embedding_strides, const int64_t* target_strides, float* loss,
                    double* gram, void* workspace, b2s_stream stream, float* mean) {
  const int C = embedding_dim + sources;
  B2S_REQUIRE(embedding_dim >= 1 && sources >= 1 && C <= B2S_DC_MAX_CHANNELS,
              "embedding_dim + sources = %d outside 2..%d", C, B2S_DC_MAX_CHANNELS);
  B2S_REQUIRE(batch >= 0 && batch <= kMaxTickets && bins >= 1 && max_frames >= 0, "bad extents");
  if (batch == 0) return B2S_OK;
  bool mean_folded = false;   // the Gram kernel folds the batch mean itself (ring / frame kernels)
  B2S_REQUIRE(embedding && target && meta && embedding_strides && target_strides && loss && gram &&
              workspace, "NULL pointer");
  const DcGrid g = dc_grid(batch, max_frames, bins, C, embedding_strides[2] == 1 && target_strides[2] == 1);
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  const Strides se{embedding_strides[0], embedding_strides[1], embedding_strides[2]};
  const Strides st{target_strides[0], target_strides[1], target_strides[2]};
  const size_t frame_smem = sizeof(float) * (2 * (frame_area(embedding_dim, (int)bins) + frame_area(sources, (int)bins)) +
                                            (bins + 31) / 32 * 32);
  static const bool no_frame = getenv("B2S_DC_NO_FRAME") != nullptr;
  if (g.tiled && !no_frame && se.c == bins && st.c == bins && frame_smem <= 100 * 1024 && max_frames < (1 << 30) &&
      bins < (1 << 20)) {
    static bool configured[64] = {};
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
      B2S_CUDA(cudaFuncSetAttribute(dc_gram_frame_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      B2S_CUDA(cudaFuncSetAttribute(dc_gram_frame_kernel<513, 20, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      configured[dev & 63] = true;
    }
    // ring form: one CTA of 12 warps per SM, four whole-frame stages (B2S_DC_RING=0: the six-warp kernel below)
    static const bool ring = [] { const char* e = getenv("B2S_DC_RING"); return e ? atoi(e) != 0 : true; }();
    const size_t ring_smem = sizeof(float) * (kRgStages * (frame_area(embedding_dim, (int)bins) + frame_area(sources, (int)bins)) +
                                             (bins + 31) / 32 * 32);
    // (the run-time-geometry instance of the ring kernel spills at its 168-register budget and measured SLOWER than
    // the six-warp kernel: 67 vs 59 us at 257 bins, 112 vs 70 us at E = 16 / K = 3 -- compile-time geometry only)
    bool reference_geometry = false;
    auto rkernel = dc_gram_ring_kernel<0, 0, 0>;
#define X(F_, E_, K_) if (bins == F_ && embedding_dim == E_ && sources == K_) { rkernel = dc_gram_ring_kernel<F_, E_, K_>; reference_geometry = true; }
    B2S_DC_GEOMETRIES(X)
#undef X
    static const bool ring_any = getenv("B2S_DC_RING_ANY") != nullptr;
    if (ring && (reference_geometry || ring_any) && ring_smem <= 200 * 1024) {
      static bool ring_configured[64] = {};
      if (!ring_configured[dev & 63]) {
        B2S_CUDA(cudaFuncSetAttribute(dc_gram_ring_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
#define X(F_, E_, K_) B2S_CUDA(cudaFuncSetAttribute(dc_gram_ring_kernel<F_, E_, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        B2S_DC_GEOMETRIES(X)
#undef X
        ring_configured[dev & 63] = true;
      }
      int rchunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(1, max_frames / 8),
                                                               (int64_t)kNumSMs / std::max<int64_t>(1, batch)));
      rchunks = std::min(rchunks, g.nchunks);
      // length-balanced chunks on a one-dimensional grid (B2S_DC_BALANCE=0: the same number of chunks per example)
      static const bool balance = [] { const char* e = getenv("B2S_DC_BALANCE"); return e ? atoi(e) != 0 : true; }();
      if (balance && batch <= kBalMaxBatch && batch <= kNumSMs) {
        rkernel<<<dim3((unsigned)kNumSMs), 32 * kRgWarps, ring_smem, (cudaStream_t)stream>>>(
            embedding, target, meta, se.t, st.t, (int)batch, kNumSMs, g.nchunks, (int)bins, embedding_dim, sources,
            partial, counters, gram, loss, mean);
      } else {
        rkernel<<<dim3((unsigned)batch, rchunks), 32 * kRgWarps, ring_smem, (cudaStream_t)stream>>>(
            embedding, target, meta, se.t, st.t, (int)batch, 0, rchunks, (int)bins, embedding_dim, sources,
            partial, counters, gram, loss, mean);
      }
      B2S_LAUNCH_CHECK("dc_gram_ring_kernel");
      return B2S_OK;
    }
    // frames are split over about two CTAs per SM (each double-buffers whole frames)
    int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(1, max_frames / 4),
                                                             (int64_t)kNumSMs * 2 / std::max<int64_t>(1, batch)));
    nchunks = std::min(nchunks, g.nchunks);   // the workspace is sized for g.nchunks partial matrices
    auto kernel = bins == 513 && embedding_dim == 20 && sources == 2 ? dc_gram_frame_kernel<513, 20, 2>
                                                                     : dc_gram_frame_kernel<0, 0, 0>;
    kernel<<<dim3((unsigned)batch, nchunks), 32 * kFrWarps, frame_smem, (cudaStream_t)stream>>>(
        embedding, target, meta, se.t, st.t, nchunks, (int)bins, embedding_dim, sources, partial, counters, gram, loss, mean);
    mean_folded = true;
  } else if (g.tiled) {
    dc_gram_tiled_kernel<<<dim3((unsigned)batch, g.nchunks), kTiledThreads, 0, (cudaStream_t)stream>>>(
        embedding, target, meta, se, st, g.nchunks, bins, embedding_dim, sources, partial, counters, gram, loss);
  } else {
    dc_gram_kernel<<<dim3((unsigned)batch, g.nchunks), 32 * g.warps, 0, (cudaStream_t)stream>>>(
        embedding, target, meta, se, st, g.nchunks, bins, embedding_dim, sources, g.blocks, g.pairs, partial,
        counters, gram, loss);
  }
  B2S_LAUNCH_CHECK("dc_gram_kernel");
  if (mean && !mean_folded) {
    dc_mean_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss, (int)batch, mean);
    B2S_LAUNCH_CHECK("dc_mean_kernel");
  }
  return B2S_OK;
}
}  // namespace

extern "C" {

int b2s_dc_forward(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                   int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                   const int64_t* embedding_strides, const int64_t* target_strides, float* loss,
                   double* gram, void* workspace, b2s_stream stream) {
  return dc_forward_impl(embedding, target, meta, batch, max_frames, bins, embedding_dim, sources, embedding_strides,
                         target_strides, loss, gram, workspace, stream, nullptr);
}

int b2s_dc_forward_mean(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                        int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                        const int64_t* embedding_strides, const int64_t* target_strides, float* loss,
                        float* mean, double* gram, void* workspace, b2s_stream stream) {
  B2S_REQUIRE(mean != nullptr && batch >= 1 && batch < kMaxTickets, "b2s_dc_forward_mean needs batch >= 1 and a mean pointer");
  return dc_forward_impl(embedding, target, meta, batch, max_frames, bins, embedding_dim, sources, embedding_strides,
                         target_strides, loss, gram, workspace, stream, mean);
}

}  // extern "C"

namespace {
int dc_backward_impl(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                     int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                     const int64_t* embedding_strides, const int64_t* target_strides, const double* gram,
                     const float* grad_loss, int64_t gl_stride, double gl_scale, float* grad_embedding,
                     b2s_stream stream) {
  const int C = embedding_dim + sources;
  B2S_REQUIRE(embedding_dim >= 1 && sources >= 1 && C <= B2S_DC_MAX_CHANNELS,
              "embedding_dim + sources = %d outside 2..%d", C, B2S_DC_MAX_CHANNELS);
  B2S_REQUIRE(batch >= 0 && batch <= kMaxTickets && bins >= 1 && max_frames >= 0, "bad extents");
  if (batch == 0) return B2S_OK;
  B2S_REQUIRE(embedding && target && meta && embedding_strides && target_strides && gram && grad_loss &&
              grad_embedding, "NULL pointer");
  const Strides se{embedding_strides[0], embedding_strides[1], embedding_strides[2]};
  const Strides st{target_strides[0], target_strides[1], target_strides[2]};
  const int64_t points = std::max<int64_t>(1, max_frames * bins);
  const size_t frame_smem = sizeof(float) * (2 * (frame_area(embedding_dim, (int)bins) + frame_area(sources, (int)bins)) +
                                            2 * frame_area(embedding_dim, (int)bins));
  static const bool no_frame = getenv("B2S_DC_NO_FRAME") != nullptr;
  if (!no_frame && C <= kTC && embedding_dim <= 25 && se.f == 1 && st.f == 1 && se.c == bins && st.c == bins &&
      frame_smem <= 200 * 1024 && max_frames < (1 << 30) && bins < (1 << 20)) {
    static bool configured[64] = {};
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
      B2S_CUDA(cudaFuncSetAttribute(dc_backward_frame_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
#define X(F_, E_, K_) B2S_CUDA(cudaFuncSetAttribute(dc_backward_frame_kernel<F_, E_, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      B2S_DC_GEOMETRIES(X)
#undef X
      configured[dev & 63] = true;
    }
    const int groups = (embedding_dim + kFbOut - 1) / kFbOut;
    const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(1, max_frames / 4),
                                                                   (int64_t)kNumSMs / std::max<int64_t>(1, batch)));
    auto kernel = dc_backward_frame_kernel<0, 0, 0>;
#define X(F_, E_, K_) if (bins == F_ && embedding_dim == E_ && sources == K_) kernel = dc_backward_frame_kernel<F_, E_, K_>;
    B2S_DC_GEOMETRIES(X)
#undef X
    static const bool balance = [] { const char* e = getenv("B2S_DC_BALANCE"); return e ? atoi(e) != 0 : true; }();
    if (balance && batch <= kBalMaxBatch && batch <= kNumSMs)
      kernel<<<dim3((unsigned)kNumSMs), 64 * groups, frame_smem, (cudaStream_t)stream>>>(
          embedding, target, meta, se.t, st.t, 0, (int)batch, kNumSMs, (int)bins, embedding_dim, sources, gram, grad_loss, gl_stride, gl_scale,
          grad_embedding);
    else
      kernel<<<dim3((unsigned)batch, nchunks), 64 * groups, frame_smem, (cudaStream_t)stream>>>(
          embedding, target, meta, se.t, st.t, nchunks, (int)batch, 0, (int)bins, embedding_dim, sources, gram, grad_loss, gl_stride, gl_scale,
          grad_embedding);
    B2S_LAUNCH_CHECK("dc_backward_frame_kernel");
    return B2S_OK;
  }
  if (C <= kTC && embedding_dim <= kBwdGroups * kBwdOut && se.f == 1 && st.f == 1) {
    const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(points / 1024,
                                               std::max<int64_t>(1, (int64_t)kNumSMs * 2 * 4 / batch)));
    dc_backward_tiled_kernel<<<dim3((unsigned)batch, nchunks), kTiledThreads, 0, (cudaStream_t)stream>>>(
        embedding, target, meta, se, st, nchunks, bins, embedding_dim, sources, gram, grad_loss, gl_stride, gl_scale, grad_embedding);
    B2S_LAUNCH_CHECK("dc_backward_tiled_kernel");
    return B2S_OK;
  }
  const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(points / 4096, 128));
  const int Ep = (embedding_dim + BS - 1) / BS * BS;
  const size_t smem = sizeof(float) * C * Ep;
  dc_backward_kernel<<<dim3((unsigned)batch, nchunks), 256, smem, (cudaStream_t)stream>>>(
      embedding, target, meta, se, st, nchunks, bins, embedding_dim, sources, gram, grad_loss, gl_stride, gl_scale,
      grad_embedding);
  B2S_LAUNCH_CHECK("dc_backward_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int b2s_dc_backward(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                    int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                    const int64_t* embedding_strides, const int64_t* target_strides, const double* gram,
                    const float* grad_loss, float* grad_embedding, b2s_stream stream) {
  return dc_backward_impl(embedding, target, meta, batch, max_frames, bins, embedding_dim, sources, embedding_strides,
                          target_strides, gram, grad_loss, 1, 1.0, grad_embedding, stream);
}

int b2s_dc_backward_scaled(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                           int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                           const int64_t* embedding_strides, const int64_t* target_strides, const double* gram,
                           const float* grad_loss, int64_t grad_loss_stride, double grad_scale,
                           float* grad_embedding, b2s_stream stream) {
  B2S_REQUIRE(grad_loss_stride == 0 || grad_loss_stride == 1, "grad_loss_stride must be 0 (broadcast) or 1");
  return dc_backward_impl(embedding, target, meta, batch, max_frames, bins, embedding_dim, sources, embedding_strides,
                          target_strides, gram, grad_loss, grad_loss_stride, grad_scale, grad_embedding, stream);
}

}  // extern "C"
