// Register-resident 1024-point real FFT / inverse real FFT, one warp per frame (sm_100a).
//
// A 1024-sample real frame is packed as 512 complex values z[n] = (x[2n], x[2n+1]); each lane keeps
// 16 of them and the warp runs a Stockham radix-8 x 8 x 8 transform with two exchanges through a
// warp-private 4 KB shared-memory tile (swizzled so every LDS.64/STS.64 is conflict free), then splits
// Z into the 513 bins of the real spectrum entirely in registers: the last pass is arranged so that a
// lane holds both Z[k] and Z[512-k].  Twiddles are per-lane constants kept in registers across frames.
//
// Butterfly assignment per lane l (j = butterfly index of a 64-butterfly pass, element r at j + 64 r):
//   passes over "natural" data (forward 1,2; inverse 2):  j in {l, l + 32}
//   passes next to the real split (forward 3; inverse 1, 3): j in {l, 64 - l}  (lane 0: {0, 32})
// For j' = 64 - l the twiddles are W8^r * conj(w_l^r): the conjugate is folded into the complex multiply
// and the W8^r factor is a cyclic shift of the radix-8 outputs, so one twiddle set serves both.
//
// The code is deliberately straight-line but compact (~900 instructions per frame): the first version
// inlined loads and epilogues into the passes, grew to 57 KB of SASS and starved on instruction fetch.
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
constexpr int kTile = 544;     // float2 slots of a warp's exchange tile (512 + padding of exchange 1)

// Packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): a complex value lives in an aligned
// register pair, so complex add / subtract is ONE instruction instead of two.  The packing moves below
// are register-allocation hints, ptxas emits no MOVs for them.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
// component-wise product (window multiply, real scaling)
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ float2 mul_i(float2 a) {
  return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// 4-point DFT with the +-i rotation of the odd difference folded into scalar adds
template <int SIGN>
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
  const float2 e0 = cadd(c0, c2), e1 = csub(c0, c2), o0 = cadd(c1, c3), d = csub(c1, c3);
  c0 = cadd(e0, o0);
  c2 = csub(e0, o0);
  if (SIGN < 0) {   // o1 = -i d = (d.y, -d.x)
    c1 = make_float2(e1.x + d.y, e1.y - d.x);
    c3 = make_float2(e1.x - d.y, e1.y + d.x);
  } else {          // o1 = +i d = (-d.y, d.x)
    c1 = make_float2(e1.x - d.y, e1.y + d.x);
    c3 = make_float2(e1.x + d.y, e1.y - d.x);
  }
}

// In-place 8-point DFT, v[q] = sum_r v[r] exp(SIGN 2 pi i r q / 8), natural order in and out.
// ROT = 1 returns the outputs cyclically shifted, v[q] <- V[(q + 1) mod 8] (see the header comment).
template <int SIGN, int ROT = 0>
__device__ __forceinline__ void radix8(float2 (&v)[8]) {
  constexpr float c = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
  float2 b1, b2, b3;
  const float2 cc = make_float2(c, c);
  if (SIGN < 0) {
    b1 = pmul(make_float2(d1.x + d1.y, d1.y - d1.x), cc);      // * (c - ic)
    b3 = pmul(make_float2(d3.y - d3.x, -(d3.x + d3.y)), cc);   // * (-c - ic)
  } else {
    b1 = pmul(make_float2(d1.x - d1.y, d1.x + d1.y), cc);      // * (c + ic)
    b3 = pmul(make_float2(-(d3.x + d3.y), d3.x - d3.y), cc);   // * (-c + ic)
  }
  b2 = mul_i<SIGN>(d2);
  dft4<SIGN>(a0, a1, a2, a3);   // even outputs 0,2,4,6
  dft4<SIGN>(b0, b1, b2, b3);   // odd outputs 1,3,5,7
  if (ROT == 0) {
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
  } else {
    v[7] = a0; v[1] = a1; v[3] = a2; v[5] = a3;
    v[0] = b0; v[2] = b1; v[4] = b2; v[6] = b3;
  }
}

// Bin held in slot p of a lane after the forward split (A side); the B side holds 512 - binA.
__device__ __forceinline__ int bin_a(int lane, int p) {
  return lane ? lane + 64 * p : (p < 4 ? 32 + 64 * p : 64 * (p - 3));
}
// lane 0 / slot 7 holds bin 256 on both sides: the B copy is a duplicate.
__device__ __forceinline__ bool bin_b_valid(int lane, int p) { return lane != 0 || p != 7; }
// packed-sample index n (z[n] = x[2n], x[2n+1]) of element r of a lane's two butterflies
__device__ __forceinline__ int natural_a(int lane, int r) { return lane + 64 * r; }
__device__ __forceinline__ int natural_b(int lane, int r) { return lane + 32 + 64 * r; }
__device__ __forceinline__ int mirrored_b(int lane, int r) { return (lane ? 64 - lane : 32) + 64 * r; }

// Per-lane constants.  `tab` = exp(-2 pi i q / 1024), q < 1024 (fp64-rounded table).
template <bool INV>
struct LaneConsts {
  float2 t2[7];   // pass 2: exp(-/+ 2 pi i (l mod 8) r / 64)
  float2 t3[7];   // pass 3: exp(-/+ 2 pi i l r / 512); lane 0 holds exp(-/+ 2 pi i r / 16) for its j = 32
  float2 ts[8];   // real-split twiddles of the 8 bin pairs (forward: times 1/2; inverse: conjugated)
  int w2e, w2o;   // exchange-2 write bases for even / odd r
  int lane;

  __device__ __forceinline__ void init(const float2* __restrict__ tab, int lane_) {
    lane = lane_;
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      float2 a = tab[((lane & 7) * r * 16) & 1023];
      float2 b = lane ? tab[(2 * lane * r) & 1023] : tab[64 * r];
      if (INV) { a.y = -a.y; b.y = -b.y; }
      t2[r - 1] = a; t3[r - 1] = b;
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float2 w = tab[bin_a(lane, p)];
      ts[p] = INV ? make_float2(w.x, -w.y) : make_float2(0.5f * w.x, 0.5f * w.y);
    }
    // exchange 1 is padded, P1(i) = i + (i >> 4); exchange 2 is swizzled, P2(i) = i ^ (((i >> 6) & 1) << 3)
    const int base2 = (lane >> 3) * 64 + (lane & 7);
    const int c = ((lane >> 3) & 1) * 8;
    w2e = base2 + c;   // even r: bit 3 of the index is clear, xor adds
    w2o = base2 - c;   // odd r:  bit 3 is set, xor subtracts
  }
};

// ---- shared-memory exchanges (warp-private kTile-float2 tile) --------------------------------------
// Exchange 1 uses the padded layout P1(i) = i + (i >> 4): element r of butterfly j (index 8 j + r) sits at
// 8 j + (j >> 1) + r and element j + 64 r at j + (j >> 4) + 68 r -- constant offsets in r on both sides,
// and conflict free for 8-byte accesses in both directions.
__device__ __forceinline__ void ex1_write(float2* tile, int j, const float2 (&v)[8]) {
  float2* p = tile + 8 * j + (j >> 1);
#pragma unroll
  for (int r = 0; r < 8; ++r) p[r] = v[r];
}
// butterflies j = l (a) and j = l + 32 (b): (j >> 4) = 0 / 1 for l < 16 and 2 / 3 above
__device__ __forceinline__ void ex1_read(const float2* tile, int lane, float2 (&a)[8], float2 (&b)[8]) {
  const float2* pa = tile + lane + (lane >> 4);
  const float2* pb = pa + 34;   // (l + 32) + ((l + 32) >> 4) = l + (l >> 4) + 34
#pragma unroll
  for (int r = 0; r < 8; ++r) { a[r] = pa[68 * r]; b[r] = pb[68 * r]; }
}
template <bool INV>
__device__ __forceinline__ void ex2_write(float2* tile, const LaneConsts<INV>& k, const float2 (&a)[8],
                                          const float2 (&b)[8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int base = (r & 1) ? k.w2o : k.w2e;
    tile[base + 8 * r] = a[r];
    tile[base + 8 * r + 256] = b[r];   // j = l + 32: (j >> 3) = (l >> 3) + 4 -> +256, same parity
  }
}
// exchange 2 read of butterfly j (< 64): swz2(j + 64 r) = (j ^ 8 (r & 1)) + 64 r
__device__ __forceinline__ void ex2_read(const float2* tile, int j, float2 (&v)[8]) {
  const int je = j, jo = j ^ 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = tile[((r & 1) ? jo : je) + 64 * r];
}

// Forward transform of windowed packed samples: a[r] = z[l + 64 r], b[r] = z[l + 32 + 64 r].
// On return slot p holds ya[p] = Y[bin_a(lane,p)], yb[p] = Y[512 - bin_a(lane,p)]; lane 0 additionally
// gets y_dc = Y[0] and y_nyq = Y[512] (both real).
__device__ __forceinline__ void rfft1024(float2 (&a)[8], float2 (&b)[8], float2* tile,
                                         const LaneConsts<false>& k, float2 (&ya)[8], float2 (&yb)[8],
                                         float& y_dc, float& y_nyq) {
  const int lane = k.lane;
  radix8<-1>(a); radix8<-1>(b);
  __syncwarp();  // previous frame's readers are done with the tile
  ex1_write(tile, lane, a);
  ex1_write(tile, lane + 32, b);
  __syncwarp();
  ex1_read(tile, lane, a, b);
#pragma unroll
  for (int r = 1; r < 8; ++r) { a[r] = cmul(a[r], k.t2[r - 1]); b[r] = cmul(b[r], k.t2[r - 1]); }
  radix8<-1>(a); radix8<-1>(b);
  __syncwarp();
  ex2_write(tile, k, a, b);
  __syncwarp();
  ex2_read(tile, lane, a);
  ex2_read(tile, lane ? 64 - lane : 32, b);
  if (lane != 0) {
#pragma unroll
    for (int r = 1; r < 8; ++r) a[r] = cmul(a[r], k.t3[r - 1]);
  }
#pragma unroll
  for (int r = 1; r < 8; ++r) b[r] = cmulc(b[r], k.t3[r - 1]);
  radix8<-1>(a); radix8<-1, 1>(b);
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
    // with Bc = conj(Z[512-k]):  Y[k] = (A+Bc)/2 + W^k (-i)(A-Bc)/2,  Y[512-k] = conj((A+Bc)/2 - W^k(-i)(A-Bc)/2)
    const float sx = A.x + B.x, sy = A.y - B.y, dx = A.x - B.x, dy = A.y + B.y;
    const float2 t = cmul(make_float2(dy, -dx), k.ts[p]);
    ya[p] = make_float2(fmaf(0.5f, sx, t.x), fmaf(0.5f, sy, t.y));
    yb[p] = make_float2(fmaf(0.5f, sx, -t.x), fmaf(-0.5f, sy, t.y));
  }
}

// Inverse: slot p carries ya[p] = Y[bin_a(lane,p)], yb[p] = Y[512 - bin_a(lane,p)] (lane 0 also y_dc,
// y_nyq).  Produces S[k] = Y0 + (-1)^k Y512 + 2 sum_{0<f<512} Re(Y_f e^{+2 pi i f k/1024}) = 1024 irfft(Y)[k]
// as packed pairs: a[r] = (S[2n], S[2n+1]) for n = natural_a(lane, r), b[r] for n = mirrored_b(lane, r).
__device__ __forceinline__ void irfft1024(const float2 (&ya)[8], const float2 (&yb)[8], float y_dc,
                                          float y_nyq, float2* tile, const LaneConsts<true>& k,
                                          float2 (&a)[8], float2 (&b)[8]) {
  const int lane = k.lane;
  const bool first = lane == 0;
  {
    float2 za[8], zb[8];  // Z'[k] (A side) and Z'[512-k] (B side) per slot, Z' = 2 Z
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float2 A = ya[p], B = make_float2(yb[p].x, -yb[p].y);
      const float2 s = cadd(A, B), d = cmul(csub(A, B), k.ts[p]);
      za[p] = make_float2(s.x - d.y, s.y + d.x);
      zb[p] = make_float2(s.x + d.y, d.x - s.y);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      // lane >= 1: a[r] = Z[lane + 64 r] = za[r];  b[r] = Z[64 - lane + 64 r] = zb[7 - r]
      // lane 0:    a[r] = Z[64 r]: r = 0 dc, r = 1..4 za[r + 3], r = 5..7 zb[11 - r];  b[r] = Z[32 + 64 r]
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
  }
  radix8<1>(a); radix8<1>(b);
  __syncwarp();
  ex1_write(tile, lane, a);
  ex1_write(tile, lane ? 64 - lane : 32, b);
  __syncwarp();
  ex1_read(tile, lane, a, b);
#pragma unroll
  for (int r = 1; r < 8; ++r) { a[r] = cmul(a[r], k.t2[r - 1]); b[r] = cmul(b[r], k.t2[r - 1]); }
  radix8<1>(a); radix8<1>(b);
  __syncwarp();
  ex2_write(tile, k, a, b);
  __syncwarp();
  ex2_read(tile, lane, a);
  ex2_read(tile, lane ? 64 - lane : 32, b);
  if (lane != 0) {
#pragma unroll
    for (int r = 1; r < 8; ++r) a[r] = cmul(a[r], k.t3[r - 1]);
  }
#pragma unroll
  for (int r = 1; r < 8; ++r) b[r] = cmulc(b[r], k.t3[r - 1]);
  radix8<1>(a); radix8<1, 1>(b);
  __syncwarp();  // the caller may now overwrite the tile
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // MUFU.SQRT, <= 2 ulp: far inside the 1e-4 budget
  return y;
}


// Loads the 16 packed sample pairs of one frame (a[r] = z[l + 64 r], b[r] = z[l + 32 + 64 r]); samples
// outside [0, samples) or beyond the window length read as zero.  `xr + s0` is frame sample 0.
template <bool VEC>
__device__ __forceinline__ void load_frame(const float* __restrict__ xr, int64_t s0, int64_t samples,
                                           int wlen, int lane, float2 (&a)[8], float2 (&b)[8]) {
  const float* base = xr + s0;
  if (s0 >= 0 && s0 + kSize <= samples && wlen == kSize) {
    if (VEC) {
      const float2* p = reinterpret_cast<const float2*>(base) + lane;
#pragma unroll
      for (int r = 0; r < 8; ++r) { a[r] = __ldg(p + 64 * r); b[r] = __ldg(p + 64 * r + 32); }
    } else {
      const float* p = base + 2 * lane;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        a[r] = make_float2(__ldg(p + 128 * r), __ldg(p + 128 * r + 1));
        b[r] = make_float2(__ldg(p + 128 * r + 64), __ldg(p + 128 * r + 65));
      }
    }
  } else {
    // valid frame-sample range [lo, hi)
    const int lo = s0 < 0 ? (int)min((int64_t)kSize, -s0) : 0;
    const int64_t room = samples - s0;
    const int hi = room <= 0 ? 0 : (int)min((int64_t)wlen, room);
    const float* p = base + 2 * lane;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int ka = 2 * lane + 128 * r, kb = ka + 64;
      a[r].x = (ka >= lo && ka < hi) ? __ldg(p + 128 * r) : 0.f;
      a[r].y = (ka + 1 >= lo && ka + 1 < hi) ? __ldg(p + 128 * r + 1) : 0.f;
      b[r].x = (kb >= lo && kb < hi) ? __ldg(p + 128 * r + 64) : 0.f;
      b[r].y = (kb + 1 >= lo && kb + 1 < hi) ? __ldg(p + 128 * r + 65) : 0.f;
    }
  }
}

// ---- asynchronous staging of signal samples (cp.async with zero fill) ------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
// 4-byte copy, zero filled when src_bytes == 0
__device__ __forceinline__ void cp_async_4_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// issue the copies of the group whose first frame starts at signal index s0 (multiple of 4)
__device__ __forceinline__ void stage_group(float* buf, const float* __restrict__ xr, int64_t s0, int span,
                                            int64_t samples) {
  // all per-chunk arithmetic in 32 bits, relative to s0
  const int c_lo = s0 < 0 ? (int)min((int64_t)span, -s0) >> 2 : 0;             // chunks before sample 0
  const int64_t room = samples - s0;
  const int n_valid = room <= 0 ? 0 : (int)min((int64_t)span, room);           // samples available from s0
  const float* base = xr + s0;
  for (int c = threadIdx.x; c < (span >> 2); c += blockDim.x) {
    int valid = min(16, 4 * (n_valid - 4 * c));                                // bytes to read, rest zero filled
    valid = (c < c_lo || valid < 0) ? 0 : valid;
    cp_async_16(buf + 4 * c, valid ? base + 4 * c : xr, valid);
  }
}

}  // namespace fft
}  // namespace b2s
