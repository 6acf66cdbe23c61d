#!/usr/bin/env python
"""Index-map prototype (numpy, CPU) of the round-2 transform: TWO real 1024-sample frames through ONE 1024-point
complex FFT on one warp, 32 values per lane.

    z[n] = w[n] (a[n] + i b[n])                       a, b: the two frames (the two sources of a position in the
                                                      fused kernel; two consecutive frames in the front-end)
    n = l + 32 p   (lane l, register p)               strided read of the staged frames: conflict free
    pass 1  (in lane):   U[l][q] = sum_p z[l + 32 p] e^{-2 pi i p q / 32}          radix-32, constant twiddles
    twiddle (per lane):  V[l][q] = U[l][q] e^{-2 pi i l q / 1024}
    exchange (shared memory, the only one): lane j receives V[l][j] for l = 0..31   (8 KB per warp and direction)
    pass 2  (in lane):   Z[j + 32 r] = sum_l V[l][j] e^{-2 pi i l r / 32}           radix-32, constant twiddles
    lane j, register r holds Z[j + 32 r];  the mirror bin 1024 - k sits in lane (32 - j) % 32, register 31 - r
    (lane 0: register (32 - r) % 32): ONE 16-value exchange between lanes j and 32 - j (shuffles), then
        A[k] = (Z[k] + conj Z[1024 - k]) / 2,   B[k] = (Z[k] - conj Z[1024 - k]) / (2 i)      additions only
    lane j ends up with bins k = j + 32 r, r = 0..15, of BOTH frames (+ bin 512 in lane 0).

Per frame: 64 shared-memory wavefronts for the exchange + 32 for the mirror shuffles, against 128 for the two
exchanges of the 8 x 8 x 8 transform of one frame as z[n] = x[2n] + i x[2n + 1]; no split twiddles.

    python tools/prototypes/pair_fft1024.py        # prints the maximum deviation from numpy.fft.rfft
"""
import numpy as np

N, LANES, REGS = 1024, 32, 32


def pair_rfft(a, b, window):
    """Spectra (513 bins each) of the two windowed real frames, computed the way the warp would."""
    z = window * (a + 1j * b)
    # registers[l][p] = z[l + 32 p]
    regs = np.array([[z[l + 32 * p] for p in range(REGS)] for l in range(LANES)])
    w32 = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    u = regs @ w32                                                        # pass 1: U[l][q], per lane
    v = u * np.exp(-2j * np.pi * np.outer(np.arange(LANES), np.arange(32)) / N)   # V[l][q]
    recv = v.T.copy()                                                     # exchange: lane j holds V[:, j]
    zk = recv @ w32                                                       # pass 2: lane j, register r: Z[j + 32 r]
    spec_a = np.zeros(N // 2 + 1, complex)
    spec_b = np.zeros(N // 2 + 1, complex)
    for j in range(LANES):
        partner = (32 - j) % 32
        for r in range(16):
            k = j + 32 * r
            mirror = zk[partner][(32 - r) % 32 if j == 0 else 31 - r]    # Z[1024 - k], from the partner lane
            spec_a[k] = 0.5 * (zk[j][r] + np.conj(mirror))
            spec_b[k] = -0.5j * (zk[j][r] - np.conj(mirror))
    # bin 512 = lane 0, register 16 (its own mirror)
    spec_a[512] = zk[0][16].real
    spec_b[512] = zk[0][16].imag
    return spec_a, spec_b


def main():
    rng = np.random.RandomState(0)
    a, b = rng.randn(N), rng.randn(N)
    from scipy.signal import windows
    window = windows.blackman(N + 1)[:-1]
    got_a, got_b = pair_rfft(a, b, window)
    want_a, want_b = np.fft.rfft(window * a), np.fft.rfft(window * b)
    err = max(np.abs(got_a - want_a).max(), np.abs(got_b - want_b).max()) / np.abs(want_a).max()
    print(f'max deviation from numpy.fft.rfft, relative to the largest bin: {err:.2e}')
    return err


if __name__ == '__main__':
    assert main() < 1e-12
