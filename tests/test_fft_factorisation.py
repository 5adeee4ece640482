"""Host-side statement of the line decompositions the pruned FFT kernels implement (csrc/fft3d.cuh), checked
against numpy: the two-factor transform of `line_phase1` / `line_phase2` and the three-factor transform of
`line3_transform`, with the factor choice of engine.cu (`factor_pair`, `factor_triple`).  The CUDA kernels
themselves are covered by the GPU tests (test_gpu_fft_paths.py: 12 grids); this file pins the index algebra -
input order j = j1*M + j2*R3 + j3, output order k = k1 + R1*(k2 + R2*k3), twiddles w_n^(m k1) and w_n^(R1 j3 k2) - and
which line lengths each variant covers."""
import numpy as np
import pytest

PAIR_RADICES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 15, 16, 18, 20]      # engine.cu factor_pair


def factor_pair(n):
    best = None
    for a in PAIR_RADICES:
        for b in PAIR_RADICES:
            if a * b == n and (best is None or max(a, b) < best[0]):
                best = (max(a, b), a, b)
    return best and best[1:]


def factor_triple(n):                                                     # engine.cu factor_triple
    if n > 420:
        return None
    best = None
    for a in range(2, 11):
        for b in range(2, a + 1):
            for c in range(2, b + 1):
                if a * b * c == n and (best is None or a < best[0]):
                    best = (a, a, b, c)
    return best and best[1:]


def smooth(n):
    for p in (2, 3, 5, 7):
        while n % p == 0:
            n //= p
    return n == 1


def line2(x, r1, r2):
    """x[j1*r2 + j2] -> X[k1 + r1*k2], unnormalised inverse DFT (exp(+2 pi i jk/n)), as in line_phase1/2."""
    n = r1 * r2
    tw = np.exp(2j * np.pi * np.arange(n) / n)
    buf = np.zeros(n, complex)
    for j2 in range(r2):
        v = np.fft.ifft(x[j2::r2]) * r1                       # r1-point DFT over j1
        for k1 in range(r1):
            buf[k1 * r2 + j2] = v[k1] * tw[j2 * k1]
    out = np.zeros(n, complex)
    for k1 in range(r1):
        v = np.fft.ifft(buf[k1 * r2:(k1 + 1) * r2]) * r2     # r2-point DFT over j2
        out[k1 + r1 * np.arange(r2)] = v
    return out


def line3(x, r1, r2, r3):
    """line3_transform: phases over j1, j2, j3 with exchange buffers A and B."""
    n, M = r1 * r2 * r3, r2 * r3
    tw = np.exp(2j * np.pi * np.arange(n) / n)
    A = np.zeros(n, complex)
    for m in range(M):
        v = np.fft.ifft(x[m::M]) * r1
        for k1 in range(r1):
            A[k1 * M + m] = v[k1] * tw[m * k1]
    B = np.zeros(n, complex)
    for k1 in range(r1):
        for j3 in range(r3):
            v = np.fft.ifft(A[k1 * M + j3 + r3 * np.arange(r2)]) * r2
            for k2 in range(r2):
                B[k1 * M + k2 * r3 + j3] = v[k2] * tw[r1 * j3 * k2]
    out = np.zeros(n, complex)
    for k1 in range(r1):
        for k2 in range(r2):
            v = np.fft.ifft(B[k1 * M + k2 * r3:k1 * M + (k2 + 1) * r3]) * r3
            out[k1 + r1 * (k2 + r2 * np.arange(r3))] = v
    return out


def test_two_factor_lines_match_numpy():
    rng = np.random.default_rng(3)
    for n in (18, 60, 90, 126, 144, 400):
        r1, r2 = factor_pair(n)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert np.abs(line2(x, r1, r2) - np.fft.ifft(x) * n).max() < 1e-11 * n


@pytest.mark.parametrize("n", [125, 147, 175, 189, 243, 245, 250, 294, 315, 336, 343, 350, 378, 384, 392, 405, 420])
def test_three_factor_lines_match_numpy(n):
    assert factor_pair(n) is None
    r1, r2, r3 = factor_triple(n)
    assert r1 * r2 * r3 == n and 10 >= r1 >= r2 >= r3 >= 2
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.abs(line3(x, r1, r2, r3) - np.fft.ifft(x) * n).max() < 1e-11 * n
    # largest twiddle indices stay inside the n-entry table
    assert (r2 * r3 - 1) * (r1 - 1) < n and r1 * (r3 - 1) * (r2 - 1) < n


def test_coverage_of_the_smooth_line_lengths():
    """Every 2-3-5-7-smooth length from 8 to 420 runs on the hand-written transform except 375 (5 x 5 x 15)."""
    missing = [n for n in range(8, 421) if smooth(n) and not factor_pair(n) and not factor_triple(n)]
    assert missing == [375]
    assert factor_pair(90) == (9, 10) and factor_pair(144) == (12, 12) and factor_pair(126) == (9, 14)
    assert factor_triple(384) == (8, 8, 6) and factor_triple(243) == (9, 9, 3) and factor_triple(432) is None
