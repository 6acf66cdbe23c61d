// The fused north-star kernel: STFT -> |.| -> mask (*) |Y| -> K x K SSE -> permutation search, with the
// target spectra |STFT(s_k)| recomputed in registers instead of being written to and re-read from HBM.
// Equals  pit_loss(mask * Y_abs[:, None, :], X_abs, axis=-2)  with  X_abs = |STFT(s)|
// (padertorch/contrib/examples/source_separation/pit/data.py:49-77, pit/model.py:117-128,
//  padertorch/ops/losses/source_separation.py:34-124).
// Algorithmic HBM bytes per utterance: 4T(1+K) + 4MFK; reading the already
// materialised |Y| instead of recomputing it trades 4T for 4MF bytes against one FFT per frame.
#include <algorithm>

#include "common.cuh"
#include "fft1024.cuh"
#include "stft_plan.cuh"
#include "perm.cuh"

using namespace b2s;


namespace {

constexpr int kFusedWarps = 4;

int fused_chunks(int64_t batch, int64_t frames) {
  const int64_t capacity = (int64_t)kNumSMs * 3;  // 3 CTAs of 4 warps per SM at ~168 registers
  int64_t c = capacity / std::max<int64_t>(1, batch);
  c = std::min<int64_t>(c, std::max<int64_t>(1, frames / kFusedWarps));
  return (int)std::max<int64_t>(1, c);
}

template <bool VEC>
__device__ __forceinline__ float2 load_pair(const float* __restrict__ xr, int64_t s0, int n, bool interior,
                                            int wlen, int64_t samples) {
  float2 v;
  if (interior) {
    if (VEC) {
      v = __ldg(reinterpret_cast<const float2*>(xr + s0) + n);
    } else {
      v.x = __ldg(xr + s0 + 2 * n);
      v.y = __ldg(xr + s0 + 2 * n + 1);
    }
  } else {
    const int64_t i0 = s0 + 2 * n, i1 = i0 + 1;
    v.x = (2 * n < wlen && i0 >= 0 && i0 < samples) ? __ldg(xr + i0) : 0.f;
    v.y = (2 * n + 1 < wlen && i1 >= 0 && i1 < samples) ? __ldg(xr + i1) : 0.f;
  }
  return v;
}

__device__ __forceinline__ float cabs2(float2 y) { return sqrtf(fmaf(y.x, y.x, y.y * y.y)); }

template <int K, bool VEC>
__global__ void __launch_bounds__(32 * kFusedWarps)
stft_pit_fused_kernel(const float* __restrict__ mixture, const float* __restrict__ yabs,
                      const float* __restrict__ sources, const float* __restrict__ mask,
                      const int64_t* __restrict__ meta, int64_t batch, int64_t samples, int64_t frames,
                      int shift, int wlen, int64_t pad_left, const float* __restrict__ win,
                      const float2* __restrict__ twtab, int nchunks, double* __restrict__ partial,
                      int* __restrict__ counters, float* __restrict__ loss, int32_t* __restrict__ perm,
                      double* __restrict__ sse) {
  constexpr int NV = K * K;
  constexpr int F = fft::kBins;
  __shared__ float2 tiles[kFusedWarps][fft::kHalf];
  __shared__ double sm[NV * kFusedWarps + NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int64_t Tb = meta ? meta[2 * b] : samples;
  const int64_t Mb = meta ? meta[2 * b + 1] : frames;
  float2* tile = tiles[warp];
  fft::LaneTwiddles<false> tw;
  tw.init(twtab, lane);
  float2 wa[8], wb[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    wa[r] = reinterpret_cast<const float2*>(win)[lane + 64 * r];
    wb[r] = reinterpret_cast<const float2*>(win)[lane + 32 + 64 * r];
  }
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  const int64_t m0 = Mb * chunk / nchunks, m1 = Mb * (chunk + 1) / nchunks;
  for (int64_t m = m0 + warp; m < m1; m += kFusedWarps) {
    const int64_t s0 = m * shift - pad_left;
    const bool interior = s0 >= 0 && s0 + wlen <= Tb && wlen == fft::kSize;
    float2 ya[8], yb[8];
    float ydc, ynyq;
    // ---- |Y| at this lane's bins
    float oa[8], ob[8], odc = 0.f, onyq = 0.f;
    if (yabs) {
      const float* row = yabs + ((int64_t)b * frames + m) * F;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int k = fft::bin_a(lane, p);
        oa[p] = __ldg(row + k);
        ob[p] = __ldg(row + fft::kHalf - k);
      }
      if (lane == 0) { odc = __ldg(row); onyq = __ldg(row + fft::kHalf); }
    } else {
      const float* xr = mixture + (int64_t)b * samples;
      auto loadz = [&](int n) -> float2 {
        const int q = n - lane, r = q >> 6;
        const float2 w = (q & 32) ? wb[r] : wa[r];
        const float2 v = load_pair<VEC>(xr, s0, n, interior, wlen, Tb);
        return make_float2(v.x * w.x, v.y * w.y);
      };
      fft::rfft1024(loadz, tile, tw, lane, ya, yb, ydc, ynyq);
#pragma unroll
      for (int p = 0; p < 8; ++p) { oa[p] = cabs2(ya[p]); ob[p] = cabs2(yb[p]); }
      odc = fabsf(ydc); onyq = fabsf(ynyq);
    }
    // ---- estimates mask_i * |Y|
    float ea[K][8], eb[K][8], edc[K], enyq[K];
    const float* mrow = mask + (((int64_t)b * frames + m) * K) * F;
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int k = fft::bin_a(lane, p);
        ea[i][p] = __ldg(mrow + i * F + k) * oa[p];
        eb[i][p] = __ldg(mrow + i * F + fft::kHalf - k) * ob[p];
      }
      edc[i] = 0.f; enyq[i] = 0.f;
      if (lane == 0) {
        edc[i] = __ldg(mrow + i * F) * odc;
        enyq[i] = __ldg(mrow + i * F + fft::kHalf) * onyq;
      }
    }
    // ---- targets: one FFT per source, folded into the SSE matrix immediately
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float* xr = sources + ((int64_t)b * K + j) * samples;
      auto loadz = [&](int n) -> float2 {
        const int q = n - lane, r = q >> 6;
        const float2 w = (q & 32) ? wb[r] : wa[r];
        const float2 v = load_pair<VEC>(xr, s0, n, interior, wlen, Tb);
        return make_float2(v.x * w.x, v.y * w.y);
      };
      fft::rfft1024(loadz, tile, tw, lane, ya, yb, ydc, ynyq);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float xa = cabs2(ya[p]);
        const float xb = fft::bin_b_valid(lane, p) ? cabs2(yb[p]) : 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float da = ea[i][p] - xa;
          const float db = fft::bin_b_valid(lane, p) ? eb[i][p] - xb : 0.f;
          acc[i * K + j] = fmaf(da, da, acc[i * K + j]);
          acc[i * K + j] = fmaf(db, db, acc[i * K + j]);
        }
      }
      if (lane == 0) {
        const float xdc = fabsf(ydc), xnyq = fabsf(ynyq);
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float d0 = edc[i] - xdc, d1 = enyq[i] - xnyq;
          acc[i * K + j] = fmaf(d0, d0, acc[i * K + j]);
          acc[i * K + j] = fmaf(d1, d1, acc[i * K + j]);
        }
      }
    }
  }

  // ---- CTA reduction, ticket, permutation search (same scheme as pit.cu)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) sm[i * kFusedWarps + warp] = (double)s;
  }
  __syncthreads();
  double* mine = partial + ((int64_t)b * nchunks + chunk) * NV;
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < kFusedWarps; ++w) s += sm[threadIdx.x * kFusedWarps + w];
    mine[threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(counters + b, 1) == nchunks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* total = sm + NV * kFusedWarps;
  if (threadIdx.x < NV) {
    double s = 0.0;
    const volatile double* p = partial + (int64_t)b * nchunks * NV + threadIdx.x;
    for (int c = 0; c < nchunks; ++c) s += p[(int64_t)c * NV];
    total[threadIdx.x] = s;
    sse[(int64_t)b * NV + threadIdx.x] = s;
  }
  __syncthreads();
  double best;
  int bp[B2S_MAX_SOURCES];
  search_permutations(total, K, best, bp);
  if (threadIdx.x == 0) {
    loss[b] = (float)(best / ((double)Mb * (double)K * (double)F));
    for (int k = 0; k < K; ++k) perm[(int64_t)b * K + k] = bp[k];
    counters[b] = 0;
  }
}

template <int K>
int launch_fused(const b2s_stft_plan* plan, const float* mixture, const float* yabs, const float* sources,
                 const float* mask, const int64_t* meta, int64_t batch, int64_t samples, int64_t frames,
                 int64_t pad_left, float* loss, int32_t* perm, double* sse, void* workspace,
                 cudaStream_t stream) {
  const int nchunks = fused_chunks(batch, frames);
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  const dim3 grid((unsigned)batch, nchunks);
  auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
  const bool vec = al8(sources) && (mixture == nullptr || al8(mixture)) && samples % 2 == 0 &&
                   plan->shift % 2 == 0 && pad_left % 2 == 0;
  if (vec)
    stft_pit_fused_kernel<K, true><<<grid, 32 * kFusedWarps, 0, stream>>>(mixture, yabs, sources, mask, meta,
        batch, samples, frames, plan->shift, plan->wlen, pad_left, plan->awin, plan->tw, nchunks, partial,
        counters, loss, perm, sse);
  else
    stft_pit_fused_kernel<K, false><<<grid, 32 * kFusedWarps, 0, stream>>>(mixture, yabs, sources, mask, meta,
        batch, samples, frames, plan->shift, plan->wlen, pad_left, plan->awin, plan->tw, nchunks, partial,
        counters, loss, perm, sse);
  B2S_LAUNCH_CHECK("stft_pit_fused_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int64_t b2s_stft_pit_workspace_bytes(int64_t batch, int64_t frames, int sources) {
  if (batch <= 0 || sources <= 0) return 16;
  return kTicketBytes + (int64_t)sizeof(double) * batch * fused_chunks(batch, frames) * sources * sources + 16;
}

int b2s_stft_pit_forward(const b2s_stft_plan* plan, const float* mixture, const float* observation_abs,
                         const float* sources, const float* mask, const int64_t* meta, int64_t batch,
                         int64_t samples, int sources_k, int64_t frames, int64_t pad_left, float* loss,
                         int32_t* perm, double* sse, void* workspace, b2s_stream stream) {
  B2S_REQUIRE(plan != nullptr, "stft plan is NULL");
  B2S_REQUIRE(plan->fast, "the fused STFT->PIT kernel exists for size-1024 plans only (got size %d)",
              plan->size);
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
