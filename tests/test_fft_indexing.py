"""The index algebra of the warp FFT kernels (csrc/fft.cuh, csrc/audio_vocos.cu), restated in
numpy and checked against numpy.fft: 512-point complex FFT as 16 points per lane x 32 lanes, the
real-FFT split step of the mel kernel, and the inverse split step of the ISTFT kernel."""
import numpy as np


def bitrev5(l):
    return int("{:05b}".format(l)[::-1], 2)


def fft512_warp(z):
    regs = np.array([[z[32 * r + lane] for r in range(16)] for lane in range(32)], dtype=complex)   # [lane][r]
    regs = np.fft.fft(regs, axis=1)                                                                  # fft16 per lane
    regs *= np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(16)) / 512)                       # W_512^(lane*k1)
    for half in (16, 8, 4, 2, 1):                                                                    # DIF over lanes
        new = regs.copy()
        for lane in range(32):
            other = regs[lane ^ half]
            if lane & half == 0:
                new[lane] = regs[lane] + other
            else:
                new[lane] = (other - regs[lane]) * np.exp(-2j * np.pi * (lane & (half - 1)) / (2 * half))
        regs = new
    Z = np.zeros(512, complex)
    for lane in range(32):
        for r in range(16):
            Z[r + 16 * bitrev5(lane)] = regs[lane, r]
    return Z


def test_fft512_distribution():
    rng = np.random.default_rng(0)
    z = rng.standard_normal(512) + 1j * rng.standard_normal(512)
    assert np.abs(fft512_warp(z) - np.fft.fft(z)).max() < 1e-10


def test_real_fft_split_step():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1024)
    Z = fft512_warp(x[0::2] + 1j * x[1::2])
    k = np.arange(513)
    zk, zc = Z[k % 512], np.conj(Z[(512 - k) % 512])
    X = (zk + zc) / 2 - 1j * np.exp(-2j * np.pi * k / 1024) * (zk - zc) / 2
    assert np.abs(X - np.fft.rfft(x)).max() < 1e-10


def test_inverse_real_fft_split_step_ignores_dc_nyquist_imag():
    rng = np.random.default_rng(2)
    S = rng.standard_normal(513) + 1j * rng.standard_normal(513)
    ref = np.fft.irfft(S, 1024)
    S2 = S.copy(); S2[0] = S2[0].real; S2[512] = S2[512].real
    k = np.arange(512)
    xk, xc = S2[k], np.conj(S2[512 - k])
    Zi = (xk + xc) / 2 + 1j * ((xk - xc) / 2 * np.exp(2j * np.pi * k / 1024))
    z = np.conj(fft512_warp(np.conj(Zi))) / 512
    x = np.empty(1024); x[0::2] = z.real; x[1::2] = z.imag
    assert np.abs(x - ref).max() < 1e-12
