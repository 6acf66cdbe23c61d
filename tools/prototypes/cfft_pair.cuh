// PROTOTYPE (not linked into libb200sep): the building blocks of the next transform kernel -- TWO real
// 1024-sample frames through ONE 1024-point complex FFT on one warp, 32 values per lane, ONE shared-memory
// exchange (index maps checked in numpy by tools/prototypes/pair_fft1024.py).
// __host__ __device__ throughout: tests/host/cfft_pair_emulate.cpp runs the 32 lanes pass by pass on the CPU.
//
//   z[n] = w[n] (a[n] + i b[n]),  n = l + 32 p  (lane l, register p)
//   pass 1 (in lane)   U[l][q] = sum_p z[l + 32 p] e^{-2 pi i p q / 32}            radix32()
//   twiddle            V[l][q] = U[l][q] e^{-2 pi i l q / 1024}                    32 per-lane constants
//   exchange           lane j receives V[l][j], l = 0..31                          padded transpose, 8 KB per warp
//   pass 2 (in lane)   Z[j + 32 r] = sum_l V[l][j] e^{-2 pi i l r / 32}            radix32()
//   mirror             Z[1024 - k] lives in lane (32 - j) % 32, register 31 - r ((32 - r) % 32 in lane 0):
//                      one 16-value exchange between lanes j and 32 - j (shuffles on the device)
//   separation         A[k] = (Z[k] + conj Z[1024 - k]) / 2,  B[k] = (Z[k] - conj Z[1024 - k]) / 2i
//                      lane j: bins k = j + 32 r, r = 0..15, of both frames (+ bin 512 in lane 0)
//
// Compile-only probe for sm_100a (nvcc -O3, one stream of two frames, scalar arithmetic, constants in registers):
// 194 registers, per TWO frames 640 FADD + 268 FMUL + 264 FFMA (the 640 additions pack into 320 FADD2), 64 LDS.32
// of input, 32 STS.64 + 32 LDS.64 for the exchange, 32 SHFL: about 425 FP issue slots and 96 exchange wavefronts
// per frame against about 480 and 128 for the 8 x 8 x 8 transform in the tree.
#pragma once
#include <cuda_runtime.h>

#ifndef B2S_HD
#define B2S_HD __host__ __device__ inline
#endif

namespace b2s {
namespace cp {

constexpr int kN = 1024;
constexpr int kLanes = 32;
constexpr int kRegs = 32;
constexpr int kPitch = 33;   // float2 slots per row of the transpose: lane l writes row l, lane j reads column j

// cos / sin of 2 pi q / 32, q = 0..15 (constexpr functions: immediates after unrolling, host and device)
B2S_HD constexpr float cos32(int q) {
  constexpr float c[16] = {1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f,
                           0.195090322f, 0.f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f,
                           -0.831469612f, -0.923879533f, -0.98078528f};
  return c[q];
}
B2S_HD constexpr float sin32(int q) {
  constexpr float s[16] = {0.f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f,
                           0.98078528f, 1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f,
                           0.382683432f, 0.195090322f};
  return s[q];
}

B2S_HD constexpr int bitrev5(int q) {
  return ((q & 1) << 4) | ((q & 2) << 2) | (q & 4) | ((q & 8) >> 2) | ((q & 16) >> 4);
}

// 32-point DFT in registers (decimation in frequency, five radix-2 levels, compile-time twiddles):
// on return v[bitrev5(q)] = sum_p v_in[p] e^{-2 pi i p q / 32}.  80 butterflies; the 49 non-trivial twiddle
// multiplications use immediates, multiplications by 1 and -i cost nothing.
B2S_HD void radix32(float2 (&v)[32]) {
#pragma unroll
  for (int level = 0; level < 5; ++level) {
    const int half = 16 >> level;
#pragma unroll
    for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 a = v[base + j], b = v[base + j + half];
        v[base + j] = make_float2(a.x + b.x, a.y + b.y);
        const float dx = a.x - b.x, dy = a.y - b.y;
        const int q = j << level;   // twiddle e^{-2 pi i q / 32}
        if (q == 0) {
          v[base + j + half] = make_float2(dx, dy);
        } else if (q == 8) {        // times -i
          v[base + j + half] = make_float2(dy, -dx);
        } else {
          const float c = cos32(q), s = sin32(q);
          v[base + j + half] = make_float2(dx * c + dy * s, dy * c - dx * s);
        }
      }
    }
  }
}

// Per-lane constants: the window at the lane's 32 sample positions and the 32 inter-pass twiddles.
struct PairConsts {
  float w[32];    // window[l + 32 p]
  float2 t[32];   // e^{-2 pi i l q / 1024}, stored at bitrev5(q) like the output of pass 1
  int lane;
  // tab = e^{-2 pi i q / 1024}, q = 0..1023
  B2S_HD void init(const float2* tab, const float* window, int lane_) {
    lane = lane_;
    for (int p = 0; p < 32; ++p) w[p] = window[lane + 32 * p];
    for (int q = 0; q < 32; ++q) t[bitrev5(q)] = tab[(lane * q) & 1023];
  }
};

// pass 1: the lane's samples of both frames (conflict-free strided reads), window, radix-32, twiddle, and the
// row of the transpose: tile[lane * kPitch + q] = V[lane][q]
B2S_HD void pass1(const float* frame_a, const float* frame_b, float2* tile, const PairConsts& k) {
  float2 v[32];
#pragma unroll
  for (int p = 0; p < 32; ++p) {
    const int n = k.lane + 32 * p;
    v[p] = make_float2(k.w[p] * frame_a[n], k.w[p] * frame_b[n]);
  }
  radix32(v);
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const float2 u = v[bitrev5(q)], t = k.t[bitrev5(q)];
    tile[k.lane * kPitch + q] = make_float2(u.x * t.x - u.y * t.y, u.x * t.y + u.y * t.x);
  }
}

// pass 2: column `lane` of the transpose, radix-32; z[r] = Z[lane + 32 r] (natural order)
B2S_HD void pass2(const float2* tile, int lane, float2 (&z)[32]) {
  float2 v[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) v[l] = tile[l * kPitch + lane];
  radix32(v);
#pragma unroll
  for (int r = 0; r < 32; ++r) z[r] = v[bitrev5(r)];
}

// register of the partner lane (32 - lane) % 32 that holds the mirror bin of this lane's register r
B2S_HD constexpr int mirror_reg(int lane, int r) { return lane == 0 ? (32 - r) & 31 : 31 - r; }

// separation of the two real spectra for this lane's bins k = lane + 32 r, r = 0..15, given the partner's mirror
// values m[r] = Z[1024 - k]; additions only
B2S_HD void separate(const float2 (&z)[32], const float2 (&m)[16], float2 (&spec_a)[16], float2 (&spec_b)[16]) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    // A = (z + conj m) / 2;  B = (z - conj m) / 2i = (-i / 2) (z - conj m)
    const float sx = z[r].x + m[r].x, sy = z[r].y - m[r].y;
    const float dx = z[r].x - m[r].x, dy = z[r].y + m[r].y;
    spec_a[r] = make_float2(0.5f * sx, 0.5f * sy);
    spec_b[r] = make_float2(0.5f * dy, -0.5f * dx);
  }
}

}  // namespace cp
}  // namespace b2s
