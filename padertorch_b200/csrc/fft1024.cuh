// Register-resident 1024-point real FFT / inverse real FFT, one warp per frame (sm_100a).
//
// A 1024-sample real frame is packed as 512 complex values z[n] = (x[2n], x[2n+1]); each lane keeps
// 16 of them and the warp runs a Stockham radix-8 x 8 x 8 transform with two exchanges through a
// warp-private 4 KB shared-memory tile (swizzled so every LDS.64/STS.64 is conflict free), then splits
// Z into the 513 bins of the real spectrum entirely in registers: the last pass is arranged so that a
// lane holds both Z[k] and Z[512-k].  Twiddles are per-lane constants kept in registers across frames.
//
// Replaces the dense windowed-DFT convolution of padertorch/ops/_stft.py:156-158 (forward) and the
// transposed convolution of :248-253 (inverse): 2.1 MFLOP/frame there, ~30 kFLOP/frame here.
#pragma once
#include <cuda_runtime.h>

namespace b2s {
namespace fft {

constexpr int kSize = 1024;    // real frame length
constexpr int kHalf = 512;     // complex transform length
constexpr int kBins = 513;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ float2 mul_i(float2 a) {
  return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

template <int SIGN>
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
  float2 e0 = cadd(c0, c2), e1 = csub(c0, c2), o0 = cadd(c1, c3), o1 = mul_i<SIGN>(csub(c1, c3));
  c0 = cadd(e0, o0); c1 = cadd(e1, o1); c2 = csub(e0, o0); c3 = csub(e1, o1);
}

// In-place 8-point DFT, v[q] = sum_r v[r] exp(SIGN 2 pi i r q / 8), natural order in and out.
template <int SIGN>
__device__ __forceinline__ void radix8(float2 (&v)[8]) {
  constexpr float c = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
  float2 b1, b2, b3;
  if (SIGN < 0) {
    b1 = make_float2(c * (d1.x + d1.y), c * (d1.y - d1.x));   // * (c - ic)
    b3 = make_float2(c * (d3.y - d3.x), -c * (d3.x + d3.y));  // * (-c - ic)
  } else {
    b1 = make_float2(c * (d1.x - d1.y), c * (d1.x + d1.y));   // * (c + ic)
    b3 = make_float2(-c * (d3.x + d3.y), c * (d3.x - d3.y));  // * (-c + ic)
  }
  b2 = mul_i<SIGN>(d2);
  dft4<SIGN>(a0, a1, a2, a3);
  dft4<SIGN>(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// shared-memory swizzles of the two exchanges (element = one float2, 512 per warp tile)
__device__ __forceinline__ int swz1(int i) { return i ^ ((i >> 3) & 15); }
__device__ __forceinline__ int swz2(int i) { return i ^ (((i >> 6) & 1) << 3); }

// Bin held in slot p of a lane after the forward split (A side); the B side holds 512 - binA.
__device__ __forceinline__ int bin_a(int lane, int p) {
  return lane ? lane + 64 * p : (p < 4 ? 32 + 64 * p : 64 * (p - 3));
}
// lane 0 / slot 7 holds bin 256 on both sides: the B copy is a duplicate.
__device__ __forceinline__ bool bin_b_valid(int lane, int p) { return lane != 0 || p != 7; }

// Per-lane twiddle constants.  `tab` = exp(-2 pi i q / 1024), q < 1024 (fp64-rounded table).
template <bool INV>
struct LaneTwiddles {
  float2 t2[7];   // pass 2 (both butterflies of a lane share them)
  float2 t3a[7];  // pass 3, butterfly a
  float2 t3b[7];  // pass 3, butterfly b
  float2 ts[8];   // real-split twiddles of the 8 bin pairs (forward: times 1/2)

  __device__ __forceinline__ void init(const float2* __restrict__ tab, int lane) {
    const int jb = INV ? lane + 32 : (lane ? 64 - lane : 32);
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      float2 a = tab[((lane & 7) * r * 16) & 1023];
      float2 b = tab[(2 * lane * r) & 1023];
      float2 c = tab[(2 * (jb & 63) * r) & 1023];
      if (INV) { a.y = -a.y; b.y = -b.y; c.y = -c.y; }
      t2[r - 1] = a; t3a[r - 1] = b; t3b[r - 1] = c;
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float2 w = tab[bin_a(lane, p)];
      ts[p] = INV ? make_float2(w.x, -w.y) : make_float2(0.5f * w.x, 0.5f * w.y);
    }
  }
};

// Forward: `loadz(n)` returns the windowed packed sample pair z[n] (n < 512).  On return slot p holds
// ya[p] = Y[bin_a(lane,p)], yb[p] = Y[512 - bin_a(lane,p)]; lane 0 additionally gets y_dc = Y[0] and
// y_nyq = Y[512] (both real).  `tile` is the warp's 512-float2 shared tile.
template <typename LoadZ>
__device__ __forceinline__ void rfft1024(LoadZ&& loadz, float2* tile, const LaneTwiddles<false>& tw,
                                         int lane, float2 (&ya)[8], float2 (&yb)[8], float& y_dc,
                                         float& y_nyq) {
  float2 a[8], b[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) { a[r] = loadz(lane + 64 * r); b[r] = loadz(lane + 32 + 64 * r); }
  radix8<-1>(a); radix8<-1>(b);
  __syncwarp();  // previous frame's readers are done with the tile
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    tile[swz1(8 * lane + r)] = a[r];
    tile[swz1(8 * (lane + 32) + r)] = b[r];
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; ++r) { a[r] = tile[swz1(lane + 64 * r)]; b[r] = tile[swz1(lane + 32 + 64 * r)]; }
#pragma unroll
  for (int r = 1; r < 8; ++r) { a[r] = cmul(a[r], tw.t2[r - 1]); b[r] = cmul(b[r], tw.t2[r - 1]); }
  radix8<-1>(a); radix8<-1>(b);
  __syncwarp();
  {
    const int ja = (lane >> 3) * 64 + (lane & 7), jb = ((lane + 32) >> 3) * 64 + (lane & 7);
#pragma unroll
    for (int r = 0; r < 8; ++r) { tile[swz2(ja + 8 * r)] = a[r]; tile[swz2(jb + 8 * r)] = b[r]; }
  }
  __syncwarp();
  {
    const int ja = lane, jb = lane ? 64 - lane : 32;
#pragma unroll
    for (int r = 0; r < 8; ++r) { a[r] = tile[swz2(ja + 64 * r)]; b[r] = tile[swz2(jb + 64 * r)]; }
  }
#pragma unroll
  for (int r = 1; r < 8; ++r) { a[r] = cmul(a[r], tw.t3a[r - 1]); b[r] = cmul(b[r], tw.t3b[r - 1]); }
  radix8<-1>(a); radix8<-1>(b);
  // a[r] = Z[ja + 64 r], b[r] = Z[jb + 64 r].  Pair Z[k] with Z[512 - k].
  const bool first = lane == 0;
  y_dc = a[0].x + a[0].y;
  y_nyq = a[0].x - a[0].y;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float2 A, B;
    if (!first) { A = a[p]; B = b[7 - p]; }
    else if (p < 4) { A = b[p]; B = b[7 - p]; }
    else { A = a[p - 3]; B = a[11 - p]; }
    // B <- conj(Z[512-k]);  Y[k] = (A+B)/2 + W^k (-i)(A-B)/2,  Y[512-k] = conj((A+B)/2 - W^k(-i)(A-B)/2)
    const float sx = A.x + B.x, sy = A.y - B.y, dx = A.x - B.x, dy = A.y + B.y;
    const float2 t = cmul(make_float2(dy, -dx), tw.ts[p]);
    ya[p] = make_float2(fmaf(0.5f, sx, t.x), fmaf(0.5f, sy, t.y));
    yb[p] = make_float2(fmaf(0.5f, sx, -t.x), fmaf(-0.5f, sy, t.y));
  }
}

// Inverse: slot p carries ya[p] = Y[bin_a(lane,p)], yb[p] = Y[512 - bin_a(lane,p)] (lane 0 also y_dc,
// y_nyq).  Produces S[k] = Y0 + (-1)^k Y512 + 2 sum_{0<f<512} Re(Y_f e^{+2 pi i f k/1024}) = 1024 irfft(Y)[k]
// as packed pairs: on return a[r] = (S[2n], S[2n+1]) for n = lane + 64 r and b[r] for n = lane + 32 + 64 r.
__device__ __forceinline__ void irfft1024(const float2 (&ya)[8], const float2 (&yb)[8], float y_dc,
                                          float y_nyq, float2* tile, const LaneTwiddles<true>& tw,
                                          int lane, float2 (&a)[8], float2 (&b)[8]) {
  float2 za[8], zb[8];  // Z'[k] (A side) and Z'[512-k] (B side) per slot, Z' = 2 Z
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float2 A = ya[p], B = make_float2(yb[p].x, -yb[p].y);
    const float2 s = cadd(A, B), d = cmul(csub(A, B), tw.ts[p]);
    za[p] = make_float2(s.x - d.y, s.y + d.x);
    zb[p] = make_float2(s.x + d.y, d.x - s.y);
  }
  const bool first = lane == 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    // lane >= 1: a[r] = Z[lane + 64 r] = za[r];  b[r] = Z[64 - lane + 64 r] = zb[7 - r]
    // lane 0:    a[r] = Z[64 r]: r=0 dc, r=1..4 za[r+3], r=5..7 zb[11-r]... ; b[r] = Z[32 + 64 r]
    float2 va, vb;
    if (!first) { va = za[r]; vb = zb[7 - r]; }
    else {
      if (r == 0) va = make_float2(y_dc + y_nyq, y_dc - y_nyq);
      else if (r <= 4) va = za[r + 3];
      else va = zb[11 - r];
      vb = r < 4 ? za[r] : zb[7 - r];
    }
    a[r] = va; b[r] = vb;
  }
  radix8<1>(a); radix8<1>(b);
  __syncwarp();
  {
    const int ja = lane, jb = lane ? 64 - lane : 32;
#pragma unroll
    for (int r = 0; r < 8; ++r) { tile[swz1(8 * ja + r)] = a[r]; tile[swz1(8 * jb + r)] = b[r]; }
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; ++r) { a[r] = tile[swz1(lane + 64 * r)]; b[r] = tile[swz1(lane + 32 + 64 * r)]; }
#pragma unroll
  for (int r = 1; r < 8; ++r) { a[r] = cmul(a[r], tw.t2[r - 1]); b[r] = cmul(b[r], tw.t2[r - 1]); }
  radix8<1>(a); radix8<1>(b);
  __syncwarp();
  {
    const int ja = (lane >> 3) * 64 + (lane & 7), jb = ((lane + 32) >> 3) * 64 + (lane & 7);
#pragma unroll
    for (int r = 0; r < 8; ++r) { tile[swz2(ja + 8 * r)] = a[r]; tile[swz2(jb + 8 * r)] = b[r]; }
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; ++r) { a[r] = tile[swz2(lane + 64 * r)]; b[r] = tile[swz2(lane + 32 + 64 * r)]; }
#pragma unroll
  for (int r = 1; r < 8; ++r) { a[r] = cmul(a[r], tw.t3a[r - 1]); b[r] = cmul(b[r], tw.t3b[r - 1]); }
  radix8<1>(a); radix8<1>(b);
  __syncwarp();  // the caller may now overwrite the tile
}

}  // namespace fft
}  // namespace b2s
