"""Thread-level NumPy model of the strided-register Stockham FFT used by csrc/fft_core.cuh.

Thread j of a line (T = N/E threads) holds positions p = j + e*T (e = 0..E-1) in registers.
Stage k (radix r, Ns = product of earlier radices): butterfly b = j + m*T (m < E/r) takes
registers e = m + t*(E/r); twiddle w_{Ns*r}^{(b mod Ns)*t}; outputs go to positions
(b div Ns)*Ns*r + (b mod Ns) + u*Ns.  All but the last stage exchange through "shared memory";
the last stage's outputs land in the thread's own register layout.
"""
import numpy as np


def fft_model(x, E, radices, sign=-1):
    N = len(x)
    T = N // E
    assert np.prod(radices) == N and all(E % r == 0 for r in radices)
    regs = np.array([[x[j + e * T] for e in range(E)] for j in range(T)], dtype=complex)
    Ns = 1
    for si, r in enumerate(radices):
        last = si == len(radices) - 1
        smem = np.zeros(N, dtype=complex)
        new = np.zeros_like(regs)
        for j in range(T):
            for m in range(E // r):
                b = j + m * T
                v = np.array([regs[j, m + t * (E // r)] for t in range(r)])
                k = b % Ns
                v = v * np.exp(sign * 2j * np.pi * k * np.arange(r) / (Ns * r))
                V = np.array([sum(v[t] * np.exp(sign * 2j * np.pi * t * u / r) for t in range(r)) for u in range(r)])
                q0 = (b // Ns) * Ns * r + k
                for u in range(r):
                    q = q0 + u * Ns
                    if last:
                        # claim: q == j + (m + u*(E//r))*T
                        assert q == j + (m + u * (E // r)) * T, (q, j, m, u)
                        new[j, m + u * (E // r)] = V[u]
                    else:
                        smem[q] = V[u]
        if not last:
            for j in range(T):
                for e in range(E):
                    new[j, e] = smem[j + e * T]
        regs = new
        Ns *= r
    out = np.zeros(N, dtype=complex)
    for j in range(T):
        for e in range(E):
            out[j + e * T] = regs[j, e]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, E, rad in [(256, 16, (16, 16)), (512, 8, (8, 8, 8)), (512, 16, (16, 2, 16)), (1024, 16, (16, 4, 16)),
                      (64, 8, (8, 8)), (128, 16, (16, 8)), (128, 8, (8, 2, 8)), (32, 8, (4, 8)), (2048, 16, (16, 8, 16)),
                      (4096, 16, (16, 16, 16)), (64, 16, (4, 16)), (1024, 8, (8, 2, 8, 8))]:
        x = rng.normal(size=N) + 1j * rng.normal(size=N)
        for sign in (-1, 1):
            y = fft_model(x, E, rad, sign)
            ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * N
            err = np.abs(y - ref).max() / np.abs(ref).max()
            print(N, E, rad, sign, f"{err:.2e}")
            assert err < 1e-12
