// Batched target / feature preparation of the PIT example on the device:
//     Y_abs = |Y|,  X_abs = |X|,  cos_phase_difference = cos(angle(Y)[:, None, :] - angle(X))
// from the complex spectra of the mixture Y [B, M, F] and of the sources X [B, K, M, F] (the output of
// b2s_stft_forward on [B, T] and [B * K, T]), with X_abs / cos_phase_difference in the model's 't k f' layout
// [B, M, K, F].  Reference: pre_batch_transform, padertorch/contrib/examples/source_separation/pit/data.py:49-77
// (per example, numpy, on the data-loader's CPU workers; its results then travel over PCIe every step).
// One streaming pass: outputs are written as aligned 16-byte stores over the FLAT output arrays (rows of 513
// floats are only 4-byte aligned, and a warp store that is not 128-byte aligned costs several times an aligned
// one); the matching inputs are gathered with 8-byte loads.
// Algorithmic bytes per utterance: 8 M F (1 + K) read + 4 M F (1 + 2 K) written.
#include "common.cuh"

using namespace b2s;

namespace {

constexpr int kTargetThreads = 256;

__device__ __forceinline__ float magnitude(float2 v) { return sqrtf(fmaf(v.x, v.x, v.y * v.y)); }

// cos(angle(y) - angle(x)) with numpy's angle(0) = 0:  Re(uy conj(ux)),  u = v / |v|  or  1 for v = 0
__device__ __forceinline__ float cos_phase_difference(float2 y, float ay, float2 x, float ax) {
  const float iy = ay > 0.f ? 1.f / ay : 0.f, ix = ax > 0.f ? 1.f / ax : 0.f;
  const float uyr = ay > 0.f ? y.x * iy : 1.f, uyi = y.y * iy;
  const float uxr = ax > 0.f ? x.x * ix : 1.f, uxi = x.y * ix;
  return fminf(1.f, fmaxf(-1.f, fmaf(uyr, uxr, uyi * uxi)));
}

// blocks [0, blocks_y): |Y| over the flat [B * M * F] array; the others: |X| and the phase term over the flat
// [B * M * K * F] output arrays.  Thread = 4 consecutive output elements.
__global__ void __launch_bounds__(kTargetThreads)
pit_targets_kernel(const float2* __restrict__ spec_y, const float2* __restrict__ spec_x, int64_t batch, int K,
                   int64_t M, int F, int64_t blocks_y, float* __restrict__ y_abs, float* __restrict__ x_abs,
                   float* __restrict__ cpd) {
  if ((int64_t)blockIdx.x < blocks_y) {
    const int64_t n = batch * M * F;
    const int64_t i0 = ((int64_t)blockIdx.x * kTargetThreads + threadIdx.x) * 4;
    if (i0 >= n) return;
    if (i0 + 4 <= n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(spec_y + i0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(spec_y + i0 + 2));
      float4 o;
      o.x = magnitude(make_float2(a.x, a.y)); o.y = magnitude(make_float2(a.z, a.w));
      o.z = magnitude(make_float2(b.x, b.y)); o.w = magnitude(make_float2(b.z, b.w));
      *reinterpret_cast<float4*>(y_abs + i0) = o;
    } else {
      for (int64_t i = i0; i < n; ++i) y_abs[i] = magnitude(__ldg(spec_y + i));
    }
    return;
  }
  const int64_t n = batch * M * K * F;
  const int64_t j0 = (((int64_t)blockIdx.x - blocks_y) * kTargetThreads + threadIdx.x) * 4;
  if (j0 >= n) return;
  // (b, m, k, f) of the first element, then incrementally
  int64_t row = j0 / F;                 // (b * M + m) * K + k
  int f = (int)(j0 - row * F);
  int64_t bm = row / K;                 // b * M + m
  int k = (int)(row - bm * K);
  int64_t b = bm / M;
  int64_t m = bm - b * M;
  float xa[4], cp[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (j0 + e < n) {
      const float2 y = __ldg(spec_y + bm * F + f);
      const float2 x = __ldg(spec_x + ((b * K + k) * M + m) * F + f);
      const float ay = magnitude(y), ax = magnitude(x);
      xa[e] = ax;
      cp[e] = cos_phase_difference(y, ay, x, ax);
    } else {
      xa[e] = 0.f; cp[e] = 0.f;
    }
    if (++f == F) {
      f = 0;
      if (++k == K) {
        k = 0; ++bm;
        if (++m == M) { m = 0; ++b; }
      }
    }
  }
  if (j0 + 4 <= n) {
    *reinterpret_cast<float4*>(x_abs + j0) = make_float4(xa[0], xa[1], xa[2], xa[3]);
    *reinterpret_cast<float4*>(cpd + j0) = make_float4(cp[0], cp[1], cp[2], cp[3]);
  } else {
    for (int e = 0; j0 + e < n; ++e) { x_abs[j0 + e] = xa[e]; cpd[j0 + e] = cp[e]; }
  }
}


// Z[b, k, m, f] = mask[b, m, k, f] * Y[b, m, f]: the masked spectra of the evaluation path
// (padertorch/contrib/examples/source_separation/pit/evaluate.py:147-152: Z = mask * Y[:, None, :], then
// 't k f -> k t f' for the inverse transform), written in the row layout b2s_istft_forward consumes
// ([B * K] rows of [M, F] complex bins).  Thread = 2 consecutive complex outputs (one 16-byte store).
__global__ void __launch_bounds__(kTargetThreads)
mask_spectrum_kernel(const float* __restrict__ mask, const float2* __restrict__ spec_y, int64_t batch, int K,
                     int64_t M, int F, float2* __restrict__ out) {
  const int64_t n = batch * K * M * F;
  const int64_t i0 = ((int64_t)blockIdx.x * kTargetThreads + threadIdx.x) * 2;
  if (i0 >= n) return;
  float2 z[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int64_t i = i0 + e;
    if (i >= n) { z[e] = make_float2(0.f, 0.f); continue; }
    const int64_t row = i / F;            // (b * K + k) * M + m
    const int f = (int)(i - row * F);
    const int64_t bk = row / M;
    const int64_t m = row - bk * M;
    const int64_t b = bk / K;
    const int k = (int)(bk - b * K);
    const float w = __ldg(mask + ((b * M + m) * K + k) * F + f);
    const float2 y = __ldg(spec_y + (b * M + m) * F + f);
    z[e] = make_float2(w * y.x, w * y.y);
  }
  if (i0 + 2 <= n) *reinterpret_cast<float4*>(out + i0) = make_float4(z[0].x, z[0].y, z[1].x, z[1].y);
  else out[i0] = z[0];
}

}  // namespace

extern "C" {

int b2s_pit_targets(const float* spec_mixture, const float* spec_sources, int64_t batch, int sources,
                    int64_t frames, int64_t bins, float* y_abs, float* x_abs, float* cos_phase_difference,
                    b2s_stream stream) {
  B2S_REQUIRE(batch >= 0 && frames >= 0 && bins >= 1 && bins < ((int64_t)1 << 30), "bad extents");
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES, "sources=%d outside the supported range 1..%d", sources,
              B2S_MAX_SOURCES);
  if (batch * frames == 0) return B2S_OK;
  B2S_REQUIRE(spec_mixture && spec_sources && y_abs && x_abs && cos_phase_difference, "NULL device pointer");
  B2S_REQUIRE(((reinterpret_cast<uintptr_t>(spec_mixture) | reinterpret_cast<uintptr_t>(spec_sources) |
                reinterpret_cast<uintptr_t>(y_abs) | reinterpret_cast<uintptr_t>(x_abs) |
                reinterpret_cast<uintptr_t>(cos_phase_difference)) & 15) == 0,
              "b2s_pit_targets needs 16-byte aligned buffers");
  const int64_t per_block = (int64_t)kTargetThreads * 4;
  const int64_t blocks_y = ceil_div(batch * frames * bins, per_block);
  const int64_t blocks_x = ceil_div(batch * frames * sources * bins, per_block);
  B2S_REQUIRE(blocks_y + blocks_x < ((int64_t)1 << 31), "too many elements for one launch");
  pit_targets_kernel<<<(unsigned)(blocks_y + blocks_x), kTargetThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(spec_mixture), reinterpret_cast<const float2*>(spec_sources), batch, sources,
      frames, (int)bins, blocks_y, y_abs, x_abs, cos_phase_difference);
  B2S_LAUNCH_CHECK("pit_targets_kernel");
  return B2S_OK;
}

int b2s_mask_spectrum(const float* mask, const float* spec_mixture, int64_t batch, int sources, int64_t frames,
                      int64_t bins, float* masked, b2s_stream stream) {
  B2S_REQUIRE(batch >= 0 && frames >= 0 && bins >= 1 && bins < ((int64_t)1 << 30), "bad extents");
  B2S_REQUIRE(sources >= 1, "sources=%d", sources);
  if (batch * frames == 0) return B2S_OK;
  B2S_REQUIRE(mask && spec_mixture && masked, "NULL device pointer");
  B2S_REQUIRE(((reinterpret_cast<uintptr_t>(spec_mixture) & 7) | (reinterpret_cast<uintptr_t>(masked) & 15)) == 0,
              "b2s_mask_spectrum needs an 8-byte aligned spectrum and a 16-byte aligned output");
  const int64_t blocks = ceil_div(batch * sources * frames * bins, (int64_t)kTargetThreads * 2);
  B2S_REQUIRE(blocks < ((int64_t)1 << 31), "too many elements for one launch");
  mask_spectrum_kernel<<<(unsigned)blocks, kTargetThreads, 0, (cudaStream_t)stream>>>(
      mask, reinterpret_cast<const float2*>(spec_mixture), batch, sources, frames, (int)bins,
      reinterpret_cast<float2*>(masked));
  B2S_LAUNCH_CHECK("mask_spectrum_kernel");
  return B2S_OK;
}

}  // extern "C"
