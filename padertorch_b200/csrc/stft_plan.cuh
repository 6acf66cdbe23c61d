// Internal layout of the opaque b2s_stft_plan (shared by stft.cu and fused.cu).
#pragma once
#include <cuda_runtime.h>

struct b2s_stft_plan {
  int device, size, shift, wlen, bins;
  int fast;      // 1: register-resident warp FFT (size 1024), 0: table-driven DFT
  float* awin;   // [size] analysis window, zero beyond wlen
  float* swin;   // [size] synthesis window (biorthogonal / size), zero beyond wlen
  float2* tw;    // [size] exp(-2 pi i q / size)
  // fast plans only: per-lane constant tables of the packed warp FFT (rfft_packed.cuh, [19][32] float4):
  float4* lane_fwd;   // analysis window, halved        (STFT forward, fused STFT -> PIT)
  float4* lane_adj;   // synthesis window, not halved    (adjoint of the iSTFT: interior bins doubled)
  // inverse transform ([23][32] float4, rf::InvLaneConsts):
  float4* lane_inv_syn;   // synthesis window, scale 1    (iSTFT)
  float4* lane_inv_ana;   // analysis window, scale 1/2   (adjoint of the STFT)
  // pair transform (cfft_pair.cuh): inter-pass twiddles exp(-2 pi i lane q / 1024) as [q][lane] (coalesced per-lane loads)
  float2* pair_tw;
  float* awin_half;   // [size] 0.5 * analysis window: the pair transform's window constants, ready to use (no arithmetic
                      // between their loads and the first frame: the loads overlap the first copy's latency)
};
