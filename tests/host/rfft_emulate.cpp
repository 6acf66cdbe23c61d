// Host emulation of padertorch_b200/csrc/rfft_packed.cuh: the 32 lanes of a warp run pass by pass on the
// CPU (shared memory = a plain array) and the 513 bins are compared with a double-precision DFT.
// Checks the index algebra (lane <-> butterfly maps, exchange layouts, twiddles, lane-0 re-pairing,
// real split) without a GPU.  Prints "max_rel_err <value>" and exits 0 when it is below 1e-5.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../padertorch_b200/csrc/rfft_packed.cuh"

using namespace b2s::rf;

int main(int argc, char** argv) {
  const int trials = argc > 1 ? atoi(argv[1]) : 4;
  std::vector<float2> tab(1024);
  for (int q = 0; q < 1024; ++q) {
    const double ang = -2.0 * M_PI * q / 1024.0;
    tab[q] = make_float2((float)cos(ang), (float)sin(ang));
  }
  double worst = 0.0;
  srand(1234);
  for (int trial = 0; trial < trials; ++trial) {
    alignas(16) static float win[1024];
    alignas(16) static float frame[1024];
    for (int i = 0; i < 1024; ++i) {
      win[i] = trial == 0 ? 1.f : (float)(0.42 - 0.5 * cos(2 * M_PI * i / 1024.0) + 0.08 * cos(4 * M_PI * i / 1024.0));
      frame[i] = (float)rand() / RAND_MAX - 0.5f;
    }
    if (trial == 1) for (int i = 0; i < 1024; ++i) frame[i] = i == 3 ? 1.f : 0.f;   // impulse
    LaneConsts k[32];
    alignas(16) static float4 table[kConstFloat4 * 32];
    for (int l = 0; l < 32; ++l) {
      LaneConsts full;
      full.init(tab.data(), win, l);
      full.pack(table);
    }
    for (int l = 0; l < 32; ++l) {
      if (trial & 1) {
        k[l].load(table, l);                       // odd trials: through the packed table
      } else {                                     // even trials: compact constants expanded on the fly
        CompactConsts c;
        c.load(table, l);
        stage_w(c, k[l]);
        stage_t2(c, k[l]);
        stage_t3(c, k[l]);
      }
    }
    alignas(16) static float2 tile[kTile];
    for (int l = 0; l < 32; ++l) pass1(frame, tile, k[l]);
    for (int l = 0; l < 32; ++l) pass2(tile, k[l]);
    std::vector<double> re(513, 1e300), im(513, 1e300);
    std::vector<int> seen(513, 0);
    for (int l = 0; l < 32; ++l) {
      float2 ya[8], yb[8];
      float ydc, ynyq;
      pass3(tile, k[l], ya, yb, ydc, ynyq);
      for (int p = 0; p < 8; ++p) {
        const int kk = bin_a(l, p);
        re[kk] = ya[p].x; im[kk] = ya[p].y; seen[kk]++;
        re[512 - kk] = yb[p].x; im[512 - kk] = -yb[p].y; seen[512 - kk]++;
      }
      if (l == 0) { re[0] = ydc; im[0] = 0; seen[0]++; re[512] = ynyq; im[512] = 0; seen[512]++; }
    }
    double maxref = 0.0, maxerr = 0.0;
    for (int f = 0; f <= 512; ++f) {
      if (!seen[f]) { printf("bin %d never produced\n", f); return 1; }
      double sr = 0, si = 0;
      for (int n = 0; n < 1024; ++n) {
        const double v = (double)frame[n] * (double)win[n], ang = -2.0 * M_PI * (double)((f * n) % 1024) / 1024.0;
        sr += v * cos(ang); si += v * sin(ang);
      }
      maxref = fmax(maxref, hypot(sr, si));
      maxerr = fmax(maxerr, hypot(sr - re[f], si - im[f]));
    }
    worst = fmax(worst, maxerr / maxref);
  }
  // ---- inverse transform: 1024 irfft(Y) with a synthesis window, both scale conventions
  for (int trial = 0; trial < trials; ++trial) {
    const float scale = (trial & 1) ? 0.5f : 1.f;
    alignas(16) static float win[1024];
    alignas(16) static float2 tile[kTile];
    std::vector<double> yr(513), yi(513);
    for (int i = 0; i < 1024; ++i) win[i] = 0.3f + (float)rand() / RAND_MAX;
    for (int f = 0; f <= 512; ++f) { yr[f] = (double)rand() / RAND_MAX - 0.5; yi[f] = (double)rand() / RAND_MAX - 0.5; }
    InvLaneConsts k[32];
    alignas(16) static float4 table[kInvConstFloat4 * 32];
    for (int l = 0; l < 32; ++l) { InvLaneConsts full; full.init(tab.data(), win, l, scale); full.pack(table); }
    for (int l = 0; l < 32; ++l) k[l].load(table, l, scale);
    for (int f = 0; f <= 512; ++f) tile[inv_bin_pos(f)] = make_float2((float)yr[f], (float)yi[f]);
    static float2 ra[32][8], rb[32][8];
    for (int l = 0; l < 32; ++l) inv_pass1_regs(tile, k[l], ra[l], rb[l]);
    for (int l = 0; l < 32; ++l) store_ex1(tile, l, ra[l], rb[l]);
    for (int l = 0; l < 32; ++l) load_ex1(tile, l, ra[l], rb[l]);
    for (int l = 0; l < 32; ++l) pass2_regs(k[l], ra[l], rb[l]);
    for (int l = 0; l < 32; ++l) store_ex2(tile, l, ra[l], rb[l]);
    for (int l = 0; l < 32; ++l) load_ex2(tile, l, ra[l], rb[l]);
    std::vector<double> got(1024, 1e300);
    for (int l = 0; l < 32; ++l) {
      if (l != 0) pass3_twiddle_a(k[l], ra[l]);
      pass3_twiddle_b(k[l], rb[l]);
      inv_finish(k[l], ra[l], rb[l]);
      for (int p = 0; p < 8; ++p) {
        const int na = l + 64 * p, nb = inv_pos_b(l, p);
        got[2 * na] = ra[l][p].x; got[2 * na + 1] = ra[l][p].y;
        got[2 * nb] = rb[l][p].x; got[2 * nb + 1] = rb[l][p].y;
      }
    }
    double maxref = 0.0, maxerr = 0.0;
    for (int m = 0; m < 1024; ++m) {
      if (got[m] > 1e299) { printf("sample %d never produced\n", m); return 1; }
      // reference: edge bins unscaled (real parts only), interior bins times 2 * scale
      double acc = yr[0] + ((m & 1) ? -yr[512] : yr[512]);
      for (int f = 1; f < 512; ++f) {
        const double ang = 2.0 * M_PI * (double)((f * m) % 1024) / 1024.0;
        acc += 2.0 * scale * (yr[f] * cos(ang) - yi[f] * sin(ang));
      }
      acc *= win[m];
      maxref = fmax(maxref, fabs(acc));
      maxerr = fmax(maxerr, fabs(acc - got[m]));
    }
    worst = fmax(worst, maxerr / maxref);
  }
  printf("max_rel_err %.3e\n", worst);
  return worst < 1e-5 ? 0 : 1;
}
