// Internal layout of the opaque b2s_stft_plan (shared by stft.cu and fused.cu).
#pragma once
#include <cuda_runtime.h>

struct b2s_stft_plan {
  int device, size, shift, wlen, bins;
  int fast;      // 1: register-resident warp FFT (size 1024), 0: table-driven DFT
  float* awin;   // [size] analysis window, zero beyond wlen
  float* swin;   // [size] synthesis window (biorthogonal / size), zero beyond wlen
  float2* tw;    // [size] exp(-2 pi i q / size)
};
