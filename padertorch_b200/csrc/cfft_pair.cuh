// Two real 1024-sample frames through ONE 1024-point complex FFT on one warp: 32 values per lane, ONE shared-memory
// exchange, addition-only separation of the two spectra (the "pair transform"; used by the fused STFT -> PIT kernel
// for the two sources of a frame position, csrc/fused_pair.cuh).  Promoted from tools/prototypes/ in round 2; index
// maps checked in numpy by tools/prototypes/pair_fft1024.py, the whole transform is __host__ __device__ and runs
// lane by lane on the CPU in tests/host/cfft_pair_emulate.cpp.
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
// Against the 8 x 8 x 8 real transform of rfft_packed.cuh, per frame: 112 instead of 160 shared-memory wavefronts
// (input 32, exchange 64, mirror shuffles 16), one exchange phase instead of two, no per-lane split twiddles, and the
// bins of a lane are j + 32 r -- an epilogue reads its operands with unit-stride loads and immediates.
// Complex additions are packed (FADD2); twiddle multiplications use immediates and stay scalar (a packed instruction
// with a swapped operand costs twice as much).
#pragma once
#include <cuda_runtime.h>

#include "rfft_packed.cuh"

#undef B2S_PAIR_HD
#define B2S_PAIR_HD __host__ __device__ __forceinline__

namespace b2s {
namespace cp {
constexpr int kN = 1024;
constexpr int kLanes = 32;
constexpr int kRegs = 32;
constexpr int kPitch = 33;   // float2 slots per row of the transpose: lane l writes row l, lane j reads column j

// cos / sin of 2 pi q / 32, q = 0..15 (constexpr functions: immediates after unrolling, host and device)
__host__ __device__ constexpr float cos32(int q) {
  constexpr float c[16] = {1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f,
                           0.195090322f, 0.f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f,
                           -0.831469612f, -0.923879533f, -0.98078528f};
  return c[q];
}
__host__ __device__ constexpr float sin32(int q) {
  constexpr float s[16] = {0.f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f,
                           0.98078528f, 1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f,
                           0.382683432f, 0.195090322f};
  return s[q];
}

// where radix32() leaves output q (natural order; kept as a function so that callers do not depend on it)
__host__ __device__ constexpr int out_pos(int q) { return q; }

// u * e^{-2 pi i m / 32}, m a compile-time constant (0 <= m < 32): immediates, the trivial factors cost nothing
template <int M>
B2S_PAIR_HD float2 twiddle32(float2 u) {
  constexpr int m = M & 31;
  if (m == 0) return u;
  if (m == 8) return make_float2(u.y, -u.x);         // -i
  if (m == 16) return make_float2(-u.x, -u.y);
  if (m == 24) return make_float2(-u.y, u.x);        // +i
  constexpr float sign = m < 16 ? 1.f : -1.f;
  constexpr float c = sign * cos32(m & 15), s = sign * sin32(m & 15);   // e^{-2 pi i m / 32} = c - i s
  return make_float2(fmaf(u.x, c, u.y * s), fmaf(u.y, c, -u.x * s));
}

// 32-point DFT in registers as 8 x 4 (p = p1 + 4 p2, q = 8 q1 + q2):
//   V[8 q1 + q2] = sum_p1 W4^(p1 q1) [ W32^(p1 q2) sum_p2 v[p1 + 4 p2] W8^(p2 q2) ]
// four radix-8 butterflies (rf::radix8: packed additions, 52 scalar-equivalent operations each), 21 twiddles with
// immediates, eight radix-4 butterflies.  On return v[q] = sum_p v_in[p] e^{-2 pi i p q / 32} (natural order).
// 120 packed + ~180 scalar instructions (five radix-2 levels: 160 + 196).
B2S_PAIR_HD void radix32(float2 (&v)[32]) {
  float2 g[4][8];
#pragma unroll
  for (int p1 = 0; p1 < 4; ++p1) {
#pragma unroll
    for (int p2 = 0; p2 < 8; ++p2) g[p1][p2] = v[p1 + 4 * p2];
    rf::radix8(g[p1]);
  }
#define B2S_TW(P1, Q2) g[P1][Q2] = twiddle32<P1 * Q2>(g[P1][Q2]);
  B2S_TW(1, 1) B2S_TW(1, 2) B2S_TW(1, 3) B2S_TW(1, 4) B2S_TW(1, 5) B2S_TW(1, 6) B2S_TW(1, 7)
  B2S_TW(2, 1) B2S_TW(2, 2) B2S_TW(2, 3) B2S_TW(2, 4) B2S_TW(2, 5) B2S_TW(2, 6) B2S_TW(2, 7)
  B2S_TW(3, 1) B2S_TW(3, 2) B2S_TW(3, 3) B2S_TW(3, 4) B2S_TW(3, 5) B2S_TW(3, 6) B2S_TW(3, 7)
#undef B2S_TW
#pragma unroll
  for (int q2 = 0; q2 < 8; ++q2) {
    const float2 t0 = rf::add2(g[0][q2], g[2][q2]), t1 = rf::sub2(g[0][q2], g[2][q2]);
    const float2 t2 = rf::add2(g[1][q2], g[3][q2]), t3 = rf::sub2(g[1][q2], g[3][q2]);
    v[q2] = rf::add2(t0, t2);
    v[16 + q2] = rf::sub2(t0, t2);
    v[8 + q2] = rf::add_mi(t1, t3);    // t1 - i t3
    v[24 + q2] = rf::add_pi(t1, t3);   // t1 + i t3
  }
}

// Per-lane constants: the window at the lane's 32 sample positions and the 32 inter-pass twiddles.
struct PairConsts {
  float w[32];    // window[l + 32 p]
  float2 t[32];   // e^{-2 pi i l q / 1024}, stored at out_pos(q) like the output of pass 1
  int lane;
  // tab = e^{-2 pi i q / 1024}, q = 0..1023
  // scale = 1/2: the factor of the separation, folded into the window
  B2S_PAIR_HD void init(const float2* tab, const float* window, int lane_, float scale = 0.5f) {
    lane = lane_;
    for (int p = 0; p < 32; ++p) w[p] = scale * window[lane + 32 * p];
    for (int q = 0; q < 32; ++q) t[out_pos(q)] = tab[(lane * q) & 1023];
  }
};

// pass 1: the lane's samples of both frames (conflict-free strided reads), window, radix-32, twiddle, and the
// row of the transpose: tile[lane * kPitch + q] = V[lane][q]
B2S_PAIR_HD void pass1(const float* frame_a, const float* frame_b, float2* tile, const PairConsts& k) {
  float2 v[32];
#pragma unroll
  for (int p = 0; p < 32; ++p) {
    const int n = k.lane + 32 * p;
    v[p] = make_float2(k.w[p] * frame_a[n], k.w[p] * frame_b[n]);
  }
  radix32(v);
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const float2 u = v[out_pos(q)], t = k.t[out_pos(q)];
    tile[k.lane * kPitch + q] = make_float2(u.x * t.x - u.y * t.y, u.x * t.y + u.y * t.x);
  }
}

// pass 2: column `lane` of the transpose, radix-32; z[r] = Z[lane + 32 r] (natural order)
B2S_PAIR_HD void pass2(const float2* tile, int lane, float2 (&z)[32]) {
  float2 v[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) v[l] = tile[l * kPitch + lane];
  radix32(v);
#pragma unroll
  for (int r = 0; r < 32; ++r) z[r] = v[out_pos(r)];
}

// register of the partner lane (32 - lane) % 32 that holds the mirror bin of this lane's register r
__host__ __device__ constexpr int mirror_reg(int lane, int r) { return lane == 0 ? (32 - r) & 31 : 31 - r; }

// separation of the two real spectra for this lane's bins k = lane + 32 r, r = 0..15, given the partner's mirror
// values m[r] = Z[1024 - k]; additions only
B2S_PAIR_HD void separate(const float2 (&z)[32], const float2 (&m)[16], float2 (&spec_a)[16], float2 (&spec_b)[16]) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    // A = (z + conj m) / 2;  B = (z - conj m) / 2i = (-i / 2) (z - conj m)
    const float sx = z[r].x + m[r].x, sy = z[r].y - m[r].y;
    const float dx = z[r].x - m[r].x, dy = z[r].y + m[r].y;
    spec_a[r] = make_float2(sx, sy);      // (the factor 1/2 is part of the window constants)
    spec_b[r] = make_float2(dy, -dx);
  }
}

}  // namespace cp
}  // namespace b2s
