// Host emulation of padertorch_b200/csrc/cfft_pair.cuh: two real frames through one 1024-point complex FFT, the 32
// lanes of a warp run pass by pass on the CPU (shared memory = a plain array, the mirror shuffle = an array
// lookup); both spectra are compared with a double-precision DFT.  Prints "max_rel_err <value>".
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../padertorch_b200/csrc/cfft_pair.cuh"

using namespace b2s::cp;

int main() {
  std::vector<float2> tab(1024);
  for (int q = 0; q < 1024; ++q) {
    const double ang = -2.0 * M_PI * q / 1024.0;
    tab[q] = make_float2((float)cos(ang), (float)sin(ang));
  }
  double worst = 0.0;
  srand(4321);
  for (int trial = 0; trial < 4; ++trial) {
    static float win[1024], fa[1024], fb[1024];
    for (int i = 0; i < 1024; ++i) {
      win[i] = trial == 0 ? 1.f : (float)(0.42 - 0.5 * cos(2 * M_PI * i / 1024.0) + 0.08 * cos(4 * M_PI * i / 1024.0));
      fa[i] = (float)rand() / RAND_MAX - 0.5f;
      fb[i] = (float)rand() / RAND_MAX - 0.5f;
    }
    if (trial == 1) for (int i = 0; i < 1024; ++i) { fa[i] = i == 5 ? 1.f : 0.f; fb[i] = i == 700 ? -1.f : 0.f; }
    static PairConsts k[32];
    for (int l = 0; l < 32; ++l) k[l].init(tab.data(), win, l);
    static float2 tile[32 * kPitch];
    for (int l = 0; l < 32; ++l) pass1(fa, fb, tile, k[l]);
    static float2 z[32][32];
    for (int l = 0; l < 32; ++l) pass2(tile, l, z[l]);
    std::vector<double> are(513, 1e300), aim(513), bre(513, 1e300), bim(513);
    for (int l = 0; l < 32; ++l) {
      const int partner = (32 - l) & 31;
      float2 m[16], sa[16], sb[16];
      for (int r = 0; r < 16; ++r) m[r] = z[partner][mirror_reg(l, r)];   // the shuffle
      separate(z[l], m, sa, sb);
      for (int r = 0; r < 16; ++r) {
        const int bin = l + 32 * r;
        are[bin] = sa[r].x; aim[bin] = sa[r].y; bre[bin] = sb[r].x; bim[bin] = sb[r].y;
      }
      if (l == 0) { are[512] = 2 * z[0][16].x; aim[512] = 0; bre[512] = 2 * z[0][16].y; bim[512] = 0; }   // (window halved)
    }
    double maxref = 0.0, maxerr = 0.0;
    for (int f = 0; f <= 512; ++f) {
      if (are[f] > 1e299 || bre[f] > 1e299) { printf("bin %d never produced\n", f); return 1; }
      double ar = 0, ai = 0, br = 0, bi = 0;
      for (int n = 0; n < 1024; ++n) {
        const double ang = -2.0 * M_PI * (double)((f * n) % 1024) / 1024.0, c = cos(ang), s = sin(ang);
        ar += (double)fa[n] * win[n] * c; ai += (double)fa[n] * win[n] * s;
        br += (double)fb[n] * win[n] * c; bi += (double)fb[n] * win[n] * s;
      }
      maxref = fmax(maxref, fmax(hypot(ar, ai), hypot(br, bi)));
      maxerr = fmax(maxerr, fmax(hypot(ar - are[f], ai - aim[f]), hypot(br - bre[f], bi - bim[f])));
    }
    worst = fmax(worst, maxerr / maxref);
  }
  printf("max_rel_err %.3e\n", worst);
  return worst < 1e-5 ? 0 : 1;
}
