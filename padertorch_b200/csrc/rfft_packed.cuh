// Forward 1024-point real FFT, one warp per frame, written for Blackwell's packed fp32x2 pipe (sm_100a).
//
// Every complex value lives in an aligned register pair (re, im).  Complex additions, conjugations and real
// scalings are single FADD2 / FMUL2 / FFMA2 instructions; ptxas folds plain component negations in the source
// into the operand modifiers of the packed instructions (R.F32x2.HI_LO.NP = per-half negation, R.F32 = 32-bit
// broadcast).  Measured on B200 (tools/ubench/pipe_rate.cu, cycles per warp instruction per SM sub-partition):
//     FADD2 / FFMA2, plain, negated or per-half negated operands      2.0
//     FFMA2 with a broadcast (.F32) operand                            2.4
//     any packed instruction with a SWAPPED operand (R.F32x2.LO_HI)    4.2   <- twice the price
//     scalar FADD / FFMA                                               0.9 - 1.1
// i.e. packed arithmetic has the FLOP rate of scalar arithmetic at half the issue slots, and the swap modifier
// is not free.  Everything that needs swapped halves -- multiplication by +-i folded into an addition, complex
// times complex -- is therefore written with scalar instructions, the rest packed.
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
B2S_HD float2 conj(float2 a) { return make_float2(a.x, -a.y); }
// a - i b and a + i b: scalar on purpose (a packed add with a swapped operand costs twice as much)
B2S_HD float2 add_mi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
B2S_HD float2 add_pi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }
// a * w and a * conj(w): four scalar instructions
B2S_HD float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
B2S_HD float2 cmulc(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, a.y * w.y), fmaf(a.y, w.x, -a.x * w.y));
}

// Levels 2 and 3 of a forward radix-8 butterfly: a_r = v_r + v_{r+4}, d_r = v_r - v_{r+4} (r < 4) in,
// v[q] = sum_r v_r exp(-2 pi i r q / 8) out (natural order).  10 packed + 16 scalar instructions.
B2S_HD void radix8_tail(float2 a0, float2 a1, float2 a2, float2 a3, float2 d0, float2 d1, float2 d2,
                        float2 d3, float2 (&v)[8]) {
  constexpr float c = 0.70710678118654752440f;
  // even outputs: 4-point DFT of a
  {
    const float2 e0 = add2(a0, a2), e1 = sub2(a0, a2), o0 = add2(a1, a3), o1 = sub2(a1, a3);
    v[0] = add2(e0, o0);
    v[4] = sub2(e0, o0);
    v[2] = add_mi(e1, o1);
    v[6] = add_pi(e1, o1);
  }
  // odd outputs: 4-point DFT of (d0, c u1, -i d2, -c g3), u1 = (1 - i) d1, g3 = (1 + i) d3
  {
    const float2 u1 = add_mi(d1, d1), g3 = add_pi(d3, d3);
    const float2 e0 = add_mi(d0, d2), e1 = add_pi(d0, d2);
    const float2 p = sub2(u1, g3), q = add2(u1, g3);
    v[1] = fma2(p, bcast(c), e0);
    v[5] = fma2(p, bcast(-c), e0);
    v[3] = make_float2(fmaf(c, q.y, e1.x), fmaf(-c, q.x, e1.y));    // e1 - i c q
    v[7] = make_float2(fmaf(-c, q.y, e1.x), fmaf(c, q.x, e1.y));    // e1 + i c q
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

constexpr int kConstFloat4 = 19;   // float4 per lane of a precomputed constant table (LaneConsts::pack / load)

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
  // The same constants as a table [kConstFloat4][32 lanes] of float4 (built once per plan on the host): a kernel
  // prologue is 19 coalesced LDG.128 instead of ~60 indexed loads.
  void pack(float4* table) const {
    float4* t = table + lane;
    for (int i = 0; i < 8; ++i) t[32 * i] = w[i];
    const float2 tw[14] = {t2[0], t2[1], t2[2], t2[3], t2[4], t2[5], t2[6], t3[0], t3[1], t3[2], t3[3], t3[4], t3[5], t3[6]};
    for (int i = 0; i < 7; ++i) t[32 * (8 + i)] = make_float4(tw[2 * i].x, tw[2 * i].y, tw[2 * i + 1].x, tw[2 * i + 1].y);
    for (int i = 0; i < 4; ++i) t[32 * (15 + i)] = make_float4(ts[2 * i].x, ts[2 * i].y, ts[2 * i + 1].x, ts[2 * i + 1].y);
  }
  B2S_HD void load(const float4* table, int lane_) {
    lane = lane_;
    const float4* t = table + lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = t[32 * i];
    float2 tw[14];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float4 v = t[32 * (8 + i)];
      tw[2 * i] = make_float2(v.x, v.y);
      tw[2 * i + 1] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) { t2[i] = tw[i]; t3[i] = tw[7 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = t[32 * (15 + i)];
      ts[2 * i] = make_float2(v.x, v.y);
      ts[2 * i + 1] = make_float2(v.z, v.w);
    }
  }
};

// The same constants in 48 instead of 76 registers: the window, three powers (w, w^2, w^4) of the pass-2 and
// pass-3 twiddle bases and the two bases of the split twiddles; the other powers are rebuilt when a pass needs
// them (4 + 4 + 6 complex multiplications per call, shared by all streams of the call).
struct CompactConsts {
  float4 w[8];
  float2 p2[3], p3[3];   // t2[0], t2[1], t2[3] / t3[0], t3[1], t3[3]
  float2 tsl, tsh;       // ts[p] = tsl W16^p (p < 4), tsh W16^p (p >= 4)
  int lane;
  B2S_HD void from_full(const LaneConsts& k) {
    lane = k.lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = k.w[i];
    p2[0] = k.t2[0]; p2[1] = k.t2[1]; p2[2] = k.t2[3];
    p3[0] = k.t3[0]; p3[1] = k.t3[1]; p3[2] = k.t3[3];
    tsl = k.ts[0];
    tsh = make_float2(-k.ts[4].y, k.ts[4].x);   // ts[4] = tsh W16^4 = -i tsh
  }
  B2S_HD void load(const float4* table, int lane_) {
    LaneConsts k;
    k.load(table, lane_);   // the loads of unused entries are dead code
    from_full(k);
  }
};
B2S_HD void expand_powers(const float2 (&p)[3], float2 (&t)[7]) {
  t[0] = p[0]; t[1] = p[1]; t[3] = p[2];
  t[2] = cmul(p[0], p[1]);
  t[4] = cmul(p[0], p[2]);
  t[5] = cmul(p[1], p[2]);
  t[6] = cmul(t[2], p[2]);
}
// what a pass needs, materialised into a LaneConsts whose other fields stay dead
B2S_HD void stage_w(const LaneConsts& c, LaneConsts& k) {
  k.lane = c.lane;
#pragma unroll
  for (int i = 0; i < 8; ++i) k.w[i] = c.w[i];
}
B2S_HD void stage_w(const CompactConsts& c, LaneConsts& k) {
  k.lane = c.lane;
#pragma unroll
  for (int i = 0; i < 8; ++i) k.w[i] = c.w[i];
}
B2S_HD void stage_t2(const LaneConsts& c, LaneConsts& k) {
#pragma unroll
  for (int i = 0; i < 7; ++i) k.t2[i] = c.t2[i];
}
B2S_HD void stage_t2(const CompactConsts& c, LaneConsts& k) { expand_powers(c.p2, k.t2); }
B2S_HD void stage_t3(const LaneConsts& c, LaneConsts& k) {
#pragma unroll
  for (int i = 0; i < 7; ++i) k.t3[i] = c.t3[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) k.ts[i] = c.ts[i];
}
B2S_HD void stage_t3(const CompactConsts& c, LaneConsts& k) {
  expand_powers(c.p3, k.t3);
  // W16^p = exp(-2 pi i p / 16)
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  k.ts[0] = c.tsl;
  k.ts[1] = cmul(c.tsl, make_float2(c1, -s1));
  k.ts[2] = cmul(c.tsl, make_float2(h, -h));
  k.ts[3] = cmul(c.tsl, make_float2(s1, -c1));
  k.ts[4] = make_float2(c.tsh.y, -c.tsh.x);
  k.ts[5] = cmul(c.tsh, make_float2(-s1, -c1));
  k.ts[6] = cmul(c.tsh, make_float2(-h, -h));
  k.ts[7] = cmul(c.tsh, make_float2(-c1, -s1));
}

// ---- building blocks of the three passes (one frame = one "stream") -----------------------------------
// Two tile layouts: separate regions for the two exchanges (kTile float2 per stream, exchange 2 at kTile1,
// two __syncwarp() per frame) or one shared region (kTile1 float2 per stream, exchange 2 at 0, four).
B2S_HD void load_frame(const float* frame, int lane, float4* x, int count = 8) {
  const float4* src = reinterpret_cast<const float4*>(frame) + lane;
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2)
    if (n2 < count) x[n2] = src[32 * n2];
}
// The same loads from a ring of hops: the frame's four hops of 256 samples sit at float offsets hop_off[0..3] of
// `ring` (a consecutive frame shares three of them with its predecessor: only the new hop is copied)
B2S_HD void load_frame_ring(const float* ring, const int* hop_off, int lane, float4* x) {
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2)
    x[n2] = *reinterpret_cast<const float4*>(ring + hop_off[n2 >> 1] + (n2 & 1) * 128 + 4 * lane);
}
// window + radix-8 over n2 of the lane's two neighbouring butterflies (x[n2] = z[2l + 64 n2], z[2l + 1 + 64 n2])
B2S_HD void pass1_regs(const float4* x, const LaneConsts& k, float2 (&va)[8], float2 (&vb)[8]) {
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
B2S_HD void store_ex1(float2* tile, int lane, const float2 (&va)[8], const float2 (&vb)[8]) {
  float4* dst = reinterpret_cast<float4*>(tile) + lane;
#pragma unroll
  for (int k0 = 0; k0 < 8; ++k0) dst[36 * k0] = make_float4(va[k0].x, va[k0].y, vb[k0].x, vb[k0].y);
}
B2S_HD void load_ex1(const float2* tile, int lane, float2 (&va)[8], float2 (&vb)[8]) {
  const float4* src = reinterpret_cast<const float4*>(tile) + (lane & 3) + 36 * (lane >> 2);
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1) {
    const float4 x = src[4 * n1];
    va[n1] = make_float2(x.x, x.y);
    vb[n1] = make_float2(x.z, x.w);
  }
}
template <class C>
B2S_HD void pass2_regs(const C& k, float2 (&va)[8], float2 (&vb)[8]) {
#pragma unroll
  for (int n1 = 1; n1 < 8; ++n1) {
    va[n1] = cmul(va[n1], k.t2[n1 - 1]);
    vb[n1] = cmul(vb[n1], k.t2[n1 - 1]);
  }
  radix8(va);
  radix8(vb);
}
// element (n0 = 2 (l & 3) + e, k0 = l >> 2, k1) at k0 + 8 k1 + 66 n0; `ex2` = start of the exchange-2 region
B2S_HD void store_ex2(float2* ex2, int lane, const float2 (&va)[8], const float2 (&vb)[8]) {
  float2* dst = ex2 + (lane >> 2) + 132 * (lane & 3);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    dst[8 * k1] = va[k1];
    dst[8 * k1 + 66] = vb[k1];
  }
}
B2S_HD void load_ex2(const float2* ex2, int lane, float2 (&a)[8], float2 (&b)[8]) {
  const float2* sa = ex2 + lane;
  const float2* sb = ex2 + (lane ? 64 - lane : 32);
#pragma unroll
  for (int n0 = 0; n0 < 8; ++n0) {
    a[n0] = sa[66 * n0];
    b[n0] = sb[66 * n0];
  }
}
template <class C>
B2S_HD void pass3_twiddle_a(const C& k, float2 (&a)[8]) {   // lanes >= 1 only (butterfly 0 has none)
#pragma unroll
  for (int n0 = 1; n0 < 8; ++n0) a[n0] = cmul(a[n0], k.t3[n0 - 1]);
}
// butterfly 64 - l: twiddles W8^n0 conj(w^n0); the W8^n0 factor shifts the outputs by one slot:
// b[q] = Z[(64 - l) + 64 ((q - 1) & 7)], the mirror of a[p] = Z[l + 64 p] is b[(8 - p) & 7]
template <class C>
B2S_HD void pass3_twiddle_b(const C& k, float2 (&b)[8]) {
#pragma unroll
  for (int n0 = 1; n0 < 8; ++n0) b[n0] = cmulc(b[n0], k.t3[n0 - 1]);
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
// Lane 0: a[p] = Z[64 p], b[q] = Z[32 + 64 ((q - 1) & 7)].  Slots 0..3 <- butterfly 32: (Z[32 + 64 p], Z[480 - 64 p])
// = (b[p + 1], b[(8 - p) & 7]); slots 4..7 <- butterfly 0: (Z[64 (p - 3)], Z[512 - 64 (p - 3)]) = (a[p - 3], a[11 - p]).
// As a parallel move: a0 <- b1 <- a4 and the cycle (a1 b2 a5 a2 b3 a6 a3 b4 a7 a4): 13 in-place moves,
// predicated so that the other lanes keep their registers where they are (no merge copies).
B2S_HD void repair_lane0(float2 (&a)[8], float2 (&b)[8], bool first) {
  pmov(a[0], b[1], first);
  pmov(b[1], a[4], first);
  float2 t = a[1];
  pmov(a[1], b[2], first); pmov(b[2], a[5], first); pmov(a[5], a[2], first); pmov(a[2], b[3], first);
  pmov(b[3], a[6], first); pmov(a[6], a[3], first); pmov(a[3], b[4], first); pmov(b[4], a[7], first);
  pmov(a[7], a[4], first); pmov(a[4], t, first);
}

// radix-8 over n0 + real split.  On return slot p holds ya[p] = Y[bin_a(lane, p)] and yb[p] =
// conj(Y[512 - bin_a]); y_dc / y_nyq = Y[0], Y[512] (meaningful in lane 0 only).
template <bool DOUBLE_INTERIOR = false>
B2S_HD void pass3_regs(const LaneConsts& k, float2 (&a)[8], float2 (&b)[8], float2 (&ya)[8], float2 (&yb)[8],
                       float& y_dc, float& y_nyq) {
  radix8(a);
  radix8(b);
  // Z carries the factor 1/2 (window): Y[0] = 2 (Re + Im), Y[512] = 2 (Re - Im); with the un-halved window of
  // DOUBLE_INTERIOR the edge bins are not doubled
  {
    const float s = a[0].x + a[0].y, d = a[0].x - a[0].y;
    y_dc = DOUBLE_INTERIOR ? s : s + s;
    y_nyq = DOUBLE_INTERIOR ? d : d + d;
  }
  repair_lane0(a, b, k.lane == 0);
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

// ---- single-stream passes (separate exchange regions): what the host emulation drives ----------------------
B2S_HD void pass1(const float* frame, float2* tile, const LaneConsts& k) {
  float4 x[8];
  float2 va[8], vb[8];
  load_frame(frame, k.lane, x);
  pass1_regs(x, k, va, vb);
  store_ex1(tile, k.lane, va, vb);
}
B2S_HD void pass2(float2* tile, const LaneConsts& k) {
  float2 va[8], vb[8];
  load_ex1(tile, k.lane, va, vb);
  pass2_regs(k, va, vb);
  store_ex2(tile + kTile1, k.lane, va, vb);
}
template <bool DOUBLE_INTERIOR = false>
B2S_HD void pass3(const float2* tile, const LaneConsts& k, float2 (&ya)[8], float2 (&yb)[8], float& y_dc,
                  float& y_nyq) {
  float2 a[8], b[8];
  load_ex2(tile + kTile1, k.lane, a, b);
  if (k.lane != 0) pass3_twiddle_a(k, a);
  pass3_twiddle_b(k, b);
  pass3_regs<DOUBLE_INTERIOR>(k, a, b, ya, yb, y_dc, y_nyq);
}

// ---- inverse real transform ----------------------------------------------------------------------------------
// S[m] = Y0 + (-1)^m Y512 + 2 sum_{0<f<512} Re(Y_f e^{+2 pi i f m / 1024}) = 1024 irfft(Y)[m] through the SAME
// three passes: with Z[k] = (Y[k] + conj Y[512-k]) + i e^{+2 pi i k/1024} (Y[k] - conj Y[512-k]) the 512-point
// inverse transform z[n] = sum_k Z[k] e^{+2 pi i k n / 512} equals S[2n] + i S[2n+1], and z = conj(FFT(conj Z)).
// The passes therefore run on u[k] = conj Z[k] = (A + B) + tin[k] (A - B),  A = conj Y[k], B = Y[512-k],
// tin[k] = -i e^{-2 pi i k/1024}, with k in the role of the forward transform's sample index, and the output
// lane l holds conj z[n] for n = l + 64 p (a side) and n = (64 - l) + 64 ((q - 1) & 7) (b side; lane 0: 32 + ...).
// No real split and no lane-0 re-pairing is needed on this side.  The imaginary parts of Y[0] and Y[512]
// are ignored like the reference's transposed convolution does (its sine rows are zero there).
constexpr int kInvConstFloat4 = 23;
struct InvLaneConsts {
  float2 tin[16];   // tin[2 n2 + e] for the lane's input bins k = 2 l + e + 64 n2
  float2 t2[7], t3[7];
  float2 wa[8], wb[8];   // (scale w[2n], -scale w[2n+1]) of the a / b output positions (conj z -> samples)
  float dc_fix;          // 1 / scale for the k = 0 input (DC and Nyquist are not scaled)
  int lane;

  // scale: factor applied to the interior bins' contribution (1: plain irfft * 1024; 1/2: adjoint of the STFT)
  B2S_HD void init(const float2* tab, const float* win, int lane_, float scale) {
    lane = lane_;
    dc_fix = 1.f / scale;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float2 t = tab[2 * lane + e + 64 * n2];
        tin[2 * n2 + e] = make_float2(t.y, -t.x);
      }
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      t2[r - 1] = tab[(16 * r * (lane >> 2)) & 1023];
      t3[r - 1] = lane ? tab[(2 * r * lane) & 1023] : tab[64 * r];
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int na = lane + 64 * p, nb = (lane ? 64 - lane : 32) + 64 * ((p + 7) & 7);
      wa[p] = make_float2(scale * win[2 * na], -scale * win[2 * na + 1]);
      wb[p] = make_float2(scale * win[2 * nb], -scale * win[2 * nb + 1]);
    }
  }
  void pack(float4* table) const {
    float4* t = table + lane;
    float2 all[46];
    for (int i = 0; i < 16; ++i) all[i] = tin[i];
    for (int i = 0; i < 7; ++i) { all[16 + i] = t2[i]; all[23 + i] = t3[i]; }
    for (int i = 0; i < 8; ++i) { all[30 + i] = wa[i]; all[38 + i] = wb[i]; }
    for (int i = 0; i < 23; ++i) t[32 * i] = make_float4(all[2 * i].x, all[2 * i].y, all[2 * i + 1].x, all[2 * i + 1].y);
  }
  B2S_HD void load(const float4* table, int lane_, float scale) {
    lane = lane_;
    dc_fix = 1.f / scale;
    const float4* t = table + lane;
    float2 all[46];
#pragma unroll
    for (int i = 0; i < 23; ++i) {
      const float4 v = t[32 * i];
      all[2 * i] = make_float2(v.x, v.y);
      all[2 * i + 1] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tin[i] = all[i];
#pragma unroll
    for (int i = 0; i < 7; ++i) { t2[i] = all[16 + i]; t3[i] = all[23 + i]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { wa[i] = all[30 + i]; wb[i] = all[38 + i]; }
  }
};
// output sample pair index of b-side slot q
B2S_HD int inv_pos_b(int lane, int q) { return (lane ? 64 - lane : 32) + 64 * ((q + 7) & 7); }

// Shared-memory position of bin k of a staged spectrum: even bins first, odd bins from 264 -- lanes read bins
// 2 l + e + 64 n2 and their mirrors, i.e. consecutive positions (conflict free) instead of every other one.
B2S_HD int inv_bin_pos(int k) { return (k >> 1) + (k & 1) * 264; }

// inverse split + radix-8 over n2; Y = the frame's 513 bins (re, im) in shared memory at inv_bin_pos(k), or -- NATURAL
// -- in bin order as they lie in global memory (a row landed by one bulk copy: every 8-byte load then has a lane
// stride of 16 bytes, i.e. four instead of two wavefronts, against 17 staging copies per lane and frame)
template <bool NATURAL = false>
B2S_HD void inv_pass1_regs(const float2* Y, const InvLaneConsts& k, float2 (&va)[8], float2 (&vb)[8]) {
  const int lane = k.lane;
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = 2 * lane + e + 64 * n2;
      // kk and 512 - kk have the same parity: positions lane + 32 n2 (+264) and 256 - lane - 32 n2 - e (+264)
      float2 A = conj(NATURAL ? Y[kk] : Y[lane + 32 * n2 + 264 * e]);
      float2 B = NATURAL ? Y[kHalf - kk] : Y[256 - lane - 32 * n2 - e + 264 * e];
      if (n2 == 0 && e == 0 && lane == 0) { A.y = 0.f; B.y = 0.f; }
      const float2 s = add2(A, B), d = sub2(A, B);
      float2 u = add2(s, cmul(d, k.tin[2 * n2 + e]));
      if (n2 == 0 && e == 0 && lane == 0) u = mul2(u, bcast(k.dc_fix));
      if (e == 0) va[n2] = u; else vb[n2] = u;
    }
  }
  radix8(va);
  radix8(vb);
}
// windowed samples of the a / b positions: (S[2n], S[2n+1]) w
B2S_HD void inv_finish(const InvLaneConsts& k, float2 (&a)[8], float2 (&b)[8]) {
  radix8(a);
  radix8(b);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    a[p] = mul2(a[p], k.wa[p]);
    b[p] = mul2(b[p], k.wb[p]);
  }
}

#ifdef __CUDACC__
// ---- NS frames per warp, interleaved for instruction-level parallelism -------------------------------------
// The per-lane constants (76 registers) are shared by the streams, so two frames per warp cost ~50 registers
// more, not twice as many, and every latency (LDS, the 4-deep packed pipe) is covered by the other stream.
// `tile` = NS regions of kTile1 float2 (both exchanges share a region: four __syncwarp() per call).
// CONSECUTIVE: the frames are 256 samples apart in the same staged row (shift 256): their 75 % overlap is
// loaded once (8 + 2 (NS - 1) LDS.128 instead of 8 NS).  Otherwise frame s starts at frame0 + s * stride.
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `input_consumed()` is called (by the whole warp) once every lane holds its input samples in registers: the
// staged frames may be overwritten from then on (a single-slot pipeline starts its next copy there).
// `before_pass3()` is called (by the whole warp) after the exchange-2 stores, i.e. when ~2/3 of the transform are done:
// a caller waits there for operands of its epilogue, so that their loads can be scheduled into pass 3.
template <int NS, bool CONSECUTIVE, bool DOUBLE_INTERIOR, class Consts, class Hook = NoHook, class Hook3 = NoHook>
__device__ __forceinline__ void rfft_streams(const float* frame0, int stride, float2* tile, const Consts& consts,
                                             float2 (&ya)[NS][8], float2 (&yb)[NS][8], float (&y_dc)[NS],
                                             float (&y_nyq)[NS], int ablate = 0, Hook input_consumed = Hook(),
                                             const int* hop_off = nullptr, Hook3 before_pass3 = Hook3()) {
  // ablate (kernel-tuning experiments only): 2 = skip the shared-memory exchanges, 4 = skip the arithmetic
  const int lane = consts.lane;
  LaneConsts k;   // fields are materialised right before the pass that uses them
  stage_w(consts, k);
  float2 va[NS][8], vb[NS][8];
  if (CONSECUTIVE) {
    float4 x[8 + 2 * (NS - 1)];
    const float4* src = reinterpret_cast<const float4*>(frame0) + lane;
#pragma unroll
    for (int n2 = 0; n2 < 8 + 2 * (NS - 1); ++n2) x[n2] = src[32 * n2];
    if (!(ablate & 4)) {
#pragma unroll
      for (int s = 0; s < NS; ++s) pass1_regs(x + 2 * s, k, va[s], vb[s]);
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) { va[s][i] = make_float2(x[i + 2 * s].x, x[i + 2 * s].y); vb[s][i] = make_float2(x[i + 2 * s].z, x[i + 2 * s].w); }
    }
  } else {
    float4 x[NS][8];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (hop_off) load_frame_ring(frame0 + s * stride, hop_off, lane, x[s]);   // `stride` = floats per ring
      else load_frame(frame0 + s * stride, lane, x[s]);
    }
    if (!(ablate & 4)) {
#pragma unroll
      for (int s = 0; s < NS; ++s) pass1_regs(x[s], k, va[s], vb[s]);
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) { va[s][i] = make_float2(x[s][i].x, x[s][i].y); vb[s][i] = make_float2(x[s][i].z, x[s][i].w); }
    }
  }
  __syncwarp();   // the previous call's exchange-2 reads are done
  input_consumed();
  if (!(ablate & 2)) {
#pragma unroll
    for (int s = 0; s < NS; ++s) store_ex1(tile + s * kTile1, lane, va[s], vb[s]);
  }
  __syncwarp();
  if (!(ablate & 2)) {
#pragma unroll
    for (int s = 0; s < NS; ++s) load_ex1(tile + s * kTile1, lane, va[s], vb[s]);
  }
  __syncwarp();   // exchange 2 overwrites the region
  if (!(ablate & 4)) {
    stage_t2(consts, k);
#pragma unroll
    for (int s = 0; s < NS; ++s) pass2_regs(k, va[s], vb[s]);
  }
  if (!(ablate & 2)) {
#pragma unroll
    for (int s = 0; s < NS; ++s) store_ex2(tile + s * kTile1, lane, va[s], vb[s]);
  }
  before_pass3();
  __syncwarp();
  if (!(ablate & 2)) {
#pragma unroll
    for (int s = 0; s < NS; ++s) load_ex2(tile + s * kTile1, lane, va[s], vb[s]);
  }
  if (!(ablate & 4)) {
    stage_t3(consts, k);
    if (lane != 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) pass3_twiddle_a(k, va[s]);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) pass3_twiddle_b(k, vb[s]);
#pragma unroll
    for (int s = 0; s < NS; ++s) pass3_regs<DOUBLE_INTERIOR>(k, va[s], vb[s], ya[s], yb[s], y_dc[s], y_nyq[s]);
  } else {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { ya[s][i] = va[s][i]; yb[s][i] = vb[s][i]; }
      y_dc[s] = va[s][0].x; y_nyq[s] = vb[s][0].y;
    }
  }
}

// NS inverse transforms per warp.  `tile` = NS regions of kTile1 float2; on entry region s holds the 513 bins of
// stream s (it becomes the exchange buffer once every lane has its inputs); on return a[s][p] / b[s][q] are the
// windowed sample pairs of positions lane + 64 p / inv_pos_b(lane, q).
// NATURAL: the spectrum of stream s starts at tile + s * kTile1 + spec_off[s] in bin order (see inv_pass1_regs).
template <int NS, bool NATURAL = false>
__device__ __forceinline__ void irfft_streams(float2* tile, const InvLaneConsts& k, float2 (&a)[NS][8],
                                              float2 (&b)[NS][8], const int* spec_off = nullptr) {
  const int lane = k.lane;
#pragma unroll
  for (int s = 0; s < NS; ++s) inv_pass1_regs<NATURAL>(tile + s * kTile1 + (NATURAL ? spec_off[s] : 0), k, a[s], b[s]);
  __syncwarp();   // every lane holds its bins: the regions become exchange buffers
#pragma unroll
  for (int s = 0; s < NS; ++s) store_ex1(tile + s * kTile1, lane, a[s], b[s]);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NS; ++s) load_ex1(tile + s * kTile1, lane, a[s], b[s]);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NS; ++s) pass2_regs(k, a[s], b[s]);
#pragma unroll
  for (int s = 0; s < NS; ++s) store_ex2(tile + s * kTile1, lane, a[s], b[s]);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NS; ++s) load_ex2(tile + s * kTile1, lane, a[s], b[s]);
  if (lane != 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) pass3_twiddle_a(k, a[s]);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) pass3_twiddle_b(k, b[s]);
#pragma unroll
  for (int s = 0; s < NS; ++s) inv_finish(k, a[s], b[s]);
  __syncwarp();   // all exchange-2 reads are done: the caller may overwrite the regions
}
#endif

}  // namespace rf
}  // namespace b2s
