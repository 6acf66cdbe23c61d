// Forward 1024-point real FFT, one warp per frame, written for Blackwell's packed fp32x2 pipe (sm_100a).
//
// Every complex value lives in an aligned register pair (re, im) and every arithmetic instruction of the
// transform is an FADD2 / FMUL2 / FFMA2.  What makes that possible are the operand modifiers of the packed
// instructions, which ptxas folds from plain component shuffles / negations in the source:
//     R.F32x2.LO_HI      swapped halves          (multiplication by +-i costs nothing)
//     R.F32x2.HI_LO.NP   per-half negation       (conjugation costs nothing; add operands and FFMA2's addend)
//     R.F32              32-bit register broadcast (complex * complex = FMUL2 + FFMA2, twiddles stay (re, im))
// A radix-8 butterfly is 26 packed instructions, a twiddle multiplication 2, one bin pair of the real split 6.
//
// Decomposition of the 512-point complex transform of z[n] = (x[2n], x[2n+1]), n = n0 + 8 n1 + 64 n2,
// k = k0 + 8 k1 + 64 k2  (k n = 64 n2 k0 + 8 n1 (k0 + 8 k1) + n0 k  mod 512):
//   pass 1  A[n0,n1,k0] = sum_n2 z[n] W8^(n2 k0)                       lane l: n0 in {2(l&3), 2(l&3)+1}, n1 = l>>2
//   pass 2  B[n0,k0,k1] = sum_n1 A W64^(n1 k0) W8^(n1 k1)              lane l: n0 in {2(l&3), 2(l&3)+1}, k0 = l>>2
//   pass 3  Z[k]        = sum_n0 B W512^(n0 (k0 + 8 k1)) W8^(n0 k2)    lane l: j = k0 + 8 k1 in {l, 64 - l}
// Pass 1 reads its 32 samples with eight conflict-free LDS.128 (two neighbouring butterflies per load) and
// fuses the window into the first butterfly level; exchange 1 moves float4 = two complex values both ways
// (layout n0 + 8 n1 + 72 k0), exchange 2 float2 (layout j + 66 n0): every access is conflict free and every
// address is a per-lane base plus an immediate.  The two exchanges use separate tile regions, so a frame
// needs two __syncwarp().  Pass 3 gives lane l both Z[k] and Z[512 - k] (k = l + 64 p): the real split runs
// in registers.  Lane 0 owns the self-mirrored butterflies j = 0 and j = 32 and re-pairs its registers first.
//
// Shared-memory traffic (the pipe that bounds this kernel, 128 B/clk/SM): 168 wavefronts per frame.
//
// The whole transform is __host__ __device__: tests/host/rfft_emulate.cpp runs the 32 lanes of a warp phase
// by phase on the CPU against a double-precision DFT, so the index algebra is checked without a GPU.
//
// Replaces the dense windowed-DFT convolution of padertorch/ops/_stft.py:156-158.
#pragma once
#include <cuda_runtime.h>

#define B2S_HD __host__ __device__ __forceinline__

namespace b2s {
namespace rf {

constexpr int kSize = 1024;
constexpr int kHalf = 512;
constexpr int kBins = 513;
constexpr int kTile1 = 576;              // float2 slots of exchange 1 (max index 7 + 56 + 72 * 7 = 567)
constexpr int kTile2 = 528;              // float2 slots of exchange 2 (max index 63 + 66 * 7 = 525)
constexpr int kTile = kTile1 + kTile2;   // float2 slots per warp (8832 bytes)

// ---- packed arithmetic --------------------------------------------------------------------------------
B2S_HD float2 add2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
B2S_HD float2 mul2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
B2S_HD float2 fma2(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
B2S_HD float2 neg(float2 a) { return make_float2(-a.x, -a.y); }
B2S_HD float2 sub2(float2 a, float2 b) { return add2(a, neg(b)); }
B2S_HD float2 bcast(float s) { return make_float2(s, s); }
// -i a  (operand modifier LO_HI + per-half negation)
B2S_HD float2 rotm(float2 a) { return make_float2(a.y, -a.x); }
B2S_HD float2 swap(float2 a) { return make_float2(a.y, a.x); }
B2S_HD float2 conj(float2 a) { return make_float2(a.x, -a.y); }
// a * w: FMUL2 + FFMA2
B2S_HD float2 cmul(float2 a, float2 w) {
  const float2 t = mul2(bcast(w.y), swap(a));                 // (w.y a.y, w.y a.x)
  return fma2(bcast(w.x), a, make_float2(-t.x, t.y));
}
// a * conj(w)
B2S_HD float2 cmulc(float2 a, float2 w) {
  const float2 t = mul2(bcast(w.y), swap(a));
  return fma2(bcast(w.x), a, make_float2(t.x, -t.y));
}

// Levels 2 and 3 of a forward radix-8 butterfly: a_r = v_r + v_{r+4}, d_r = v_r - v_{r+4} (r < 4) in,
// v[q] = sum_r v_r exp(-2 pi i r q / 8) out (natural order).  18 packed instructions.
B2S_HD void radix8_tail(float2 a0, float2 a1, float2 a2, float2 a3, float2 d0, float2 d1, float2 d2,
                        float2 d3, float2 (&v)[8]) {
  constexpr float c = 0.70710678118654752440f;
  // even outputs: 4-point DFT of a
  {
    const float2 e0 = add2(a0, a2), e1 = sub2(a0, a2), o0 = add2(a1, a3), o1 = sub2(a1, a3);
    v[0] = add2(e0, o0);
    v[4] = sub2(e0, o0);
    v[2] = add2(e1, rotm(o1));
    v[6] = sub2(e1, rotm(o1));
  }
  // odd outputs: 4-point DFT of (d0, c u1, -i d2, -c g3), u1 = (1 - i) d1, g3 = (1 + i) d3
  {
    const float2 u1 = add2(d1, rotm(d1)), g3 = sub2(d3, rotm(d3));
    const float2 e0 = add2(d0, rotm(d2)), e1 = sub2(d0, rotm(d2));
    const float2 p = sub2(u1, g3), q = add2(u1, g3);
    v[1] = fma2(p, bcast(c), e0);
    v[5] = fma2(p, bcast(-c), e0);
    v[3] = fma2(swap(q), make_float2(c, -c), e1);    // e1 - i c q
    v[7] = fma2(swap(q), make_float2(-c, c), e1);    // e1 + i c q
  }
}

B2S_HD void radix8(float2 (&v)[8]) {
  const float2 a0 = add2(v[0], v[4]), a1 = add2(v[1], v[5]), a2 = add2(v[2], v[6]), a3 = add2(v[3], v[7]);
  const float2 d0 = sub2(v[0], v[4]), d1 = sub2(v[1], v[5]), d2 = sub2(v[2], v[6]), d3 = sub2(v[3], v[7]);
  radix8_tail(a0, a1, a2, a3, d0, d1, d2, d3, v);
}

// Bin of slot p of a lane after the split (A side); the B side holds 512 - bin_a.  Lanes >= 1: l + 64 p.
// Lane 0: slots 0..3 carry butterfly 32 (32 + 64 p), slots 4..7 butterfly 0 (64 (p - 3)).
B2S_HD int bin_a(int lane, int p) {
  return lane ? lane + 64 * p : (p < 4 ? 32 + 64 * p : 64 * (p - 3));
}

// Per-lane constants, kept in registers across frames.  `tab` = exp(-2 pi i q / 1024), `win` = the
// (zero-extended) 1024-sample window.
struct LaneConsts {
  float4 w[8];    // scale * window of the lane's samples 4 l + 128 n2 + (0..3)   (1/2: the 1/2 of the split)
  float2 t2[7];   // pass 2: exp(-2 pi i n1 (l >> 2) / 64), n1 = 1..7
  float2 t3[7];   // pass 3: exp(-2 pi i n0 l / 512); lane 0: exp(-2 pi i n0 / 16) (its butterfly 32)
  float2 ts[8];   // split: -i exp(-2 pi i bin_a(l, p) / 1024)
  int lane;

  // scale = 1/2: plain transform; scale = 1: interior bins doubled (pass3<true>)
  B2S_HD void init(const float2* tab, const float* win, int lane_, float scale = 0.5f) {
    lane = lane_;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) {
      const float* p = win + 4 * lane + 128 * n2;
      w[n2] = make_float4(scale * p[0], scale * p[1], scale * p[2], scale * p[3]);
    }
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      t2[r - 1] = tab[(16 * r * (lane >> 2)) & 1023];
      t3[r - 1] = lane ? tab[(2 * r * lane) & 1023] : tab[64 * r];
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float2 t = tab[bin_a(lane, p)];
      ts[p] = make_float2(t.y, -t.x);
    }
  }
};

// ---- the three passes ---------------------------------------------------------------------------------
// `frame` = the frame's first sample in (16-byte aligned) staged memory; `tile` = the warp's kTile float2.
// Pass 1: load, window, radix-8 over n2, write exchange 1.
B2S_HD void pass1(const float* frame, float2* tile, const LaneConsts& k) {
  const float4* src = reinterpret_cast<const float4*>(frame) + k.lane;
  float4 x[8];
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) x[n2] = src[32 * n2];
  float2 va[8], vb[8];
  {
    float2 a[4], d[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 lo = make_float2(x[r].x, x[r].y), wl = make_float2(k.w[r].x, k.w[r].y);
      const float2 hi = mul2(make_float2(x[r + 4].x, x[r + 4].y), make_float2(k.w[r + 4].x, k.w[r + 4].y));
      a[r] = fma2(lo, wl, hi);
      d[r] = fma2(lo, wl, neg(hi));
    }
    radix8_tail(a[0], a[1], a[2], a[3], d[0], d[1], d[2], d[3], va);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 lo = make_float2(x[r].z, x[r].w), wl = make_float2(k.w[r].z, k.w[r].w);
      const float2 hi = mul2(make_float2(x[r + 4].z, x[r + 4].w), make_float2(k.w[r + 4].z, k.w[r + 4].w));
      a[r] = fma2(lo, wl, hi);
      d[r] = fma2(lo, wl, neg(hi));
    }
    radix8_tail(a[0], a[1], a[2], a[3], d[0], d[1], d[2], d[3], vb);
  }
  // element (n0 = 2 (l & 3) + e, n1 = l >> 2, k0) at n0 + 8 n1 + 72 k0 = 2 l + e + 72 k0
  float4* dst = reinterpret_cast<float4*>(tile) + k.lane;
#pragma unroll
  for (int k0 = 0; k0 < 8; ++k0) dst[36 * k0] = make_float4(va[k0].x, va[k0].y, vb[k0].x, vb[k0].y);
}

// Pass 2: read exchange 1, twiddle, radix-8 over n1, twiddle-free write of exchange 2.
B2S_HD void pass2(float2* tile, const LaneConsts& k) {
  const int a2 = k.lane & 3, k0 = k.lane >> 2;
  const float4* src = reinterpret_cast<const float4*>(tile) + a2 + 36 * k0;
  float2 va[8], vb[8];
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) {
    const float4 x = src[4 * n1];
    va[n1] = make_float2(x.x, x.y);
    vb[n1] = make_float2(x.z, x.w);
  }
#pragma unroll
  for (int n1 = 1; n1 < 8; ++n1) {
    va[n1] = cmul(va[n1], k.t2[n1 - 1]);
    vb[n1] = cmul(vb[n1], k.t2[n1 - 1]);
  }
  radix8(va);
  radix8(vb);
  // element (n0 = 2 a2 + e, k0, k1) at k0 + 8 k1 + 66 n0
  float2* dst = tile + kTile1 + k0 + 132 * a2;
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    dst[8 * k1] = va[k1];
    dst[8 * k1 + 66] = vb[k1];
  }
}

// dst <- src in lane 0 only, in place
B2S_HD void pmov(float2& dst, float2 src, bool pred) {
#ifdef __CUDA_ARCH__
  asm("{.reg .pred p; setp.ne.s32 p, %4, 0; @p mov.f32 %0, %2; @p mov.f32 %1, %3;}"
      : "+f"(dst.x), "+f"(dst.y) : "f"(src.x), "f"(src.y), "r"((int)pred));
#else
  if (pred) dst = src;
#endif
}
B2S_HD void repair_lane0(float2 (&a)[8], float2 (&b)[8], bool first) {
  pmov(a[0], b[1], first);
  pmov(b[1], a[4], first);
  float2 t = a[1];
  pmov(a[1], b[2], first); pmov(b[2], a[5], first); pmov(a[5], a[2], first); pmov(a[2], b[3], first);
  pmov(b[3], a[6], first); pmov(a[6], a[3], first); pmov(a[3], b[4], first); pmov(b[4], a[7], first);
  pmov(a[7], a[4], first); pmov(a[4], t, first);
}

// Pass 3 + real split.  On return slot p holds ya[p] = Y[bin_a(lane, p)] and yb[p] = conj(Y[512 - bin_a]);
// y_dc / y_nyq = Y[0], Y[512] (meaningful in lane 0 only).
template <bool DOUBLE_INTERIOR = false>
B2S_HD void pass3(const float2* tile, const LaneConsts& k, float2 (&ya)[8], float2 (&yb)[8], float& y_dc,
                  float& y_nyq) {
  const int lane = k.lane;
  const float2* sa = tile + kTile1 + lane;
  const float2* sb = tile + kTile1 + (lane ? 64 - lane : 32);
  float2 a[8], b[8];
#pragma unroll
  for (int n0 = 0; n0 < 8; ++n0) {
    a[n0] = sa[66 * n0];
    b[n0] = sb[66 * n0];
  }
  if (lane != 0) {
#pragma unroll
    for (int n0 = 1; n0 < 8; ++n0) a[n0] = cmul(a[n0], k.t3[n0 - 1]);
  }
  // butterfly 64 - l: twiddles W8^n0 conj(w^n0); the W8^n0 factor shifts the outputs by one slot:
  // b[q] = Z[(64 - l) + 64 ((q - 1) & 7)], the mirror of a[p] = Z[l + 64 p] is b[(8 - p) & 7]
#pragma unroll
  for (int n0 = 1; n0 < 8; ++n0) b[n0] = cmulc(b[n0], k.t3[n0 - 1]);
  radix8(a);
  radix8(b);
  // Z carries the factor 1/2 (window): Y[0] = 2 (Re + Im), Y[512] = 2 (Re - Im); with the un-halved window of
  // DOUBLE_INTERIOR the edge bins are not doubled
  {
    const float s = a[0].x + a[0].y, d = a[0].x - a[0].y;
    y_dc = DOUBLE_INTERIOR ? s : s + s;
    y_nyq = DOUBLE_INTERIOR ? d : d + d;
  }
  // Lane 0: a[p] = Z[64 p], b[q] = Z[32 + 64 ((q - 1) & 7)].  Slots 0..3 <- butterfly 32: (Z[32 + 64 p], Z[480 - 64 p])
  // = (b[p + 1], b[(8 - p) & 7]); slots 4..7 <- butterfly 0: (Z[64 (p - 3)], Z[512 - 64 (p - 3)]) = (a[p - 3], a[11 - p]).
  // As a parallel move: a0 <- b1 <- a4 and the cycle (a1 b2 a5 a2 b3 a6 a3 b4 a7 a4): 13 in-place moves,
  // predicated so that the other lanes keep their registers where they are (no merge copies).
  repair_lane0(a, b, lane == 0);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float2 A = a[p], B = b[(8 - p) & 7];
    // Y[k] = (A + conj B) + ts (A - conj B),  conj(Y[512 - k]) = (A + conj B) - ts (A - conj B)
    const float2 s = add2(A, conj(B)), d = add2(A, make_float2(-B.x, B.y));
    const float2 t = cmul(d, k.ts[p]);
    ya[p] = add2(s, t);
    yb[p] = sub2(s, t);
  }
}

}  // namespace rf
}  // namespace b2s
