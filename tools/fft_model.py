"""Numpy model of the exact index / twiddle scheme used by dynpy_b200/csrc/fft_engine.cuh.

Not product code: it documents and checks the decomposition before it is transcribed
to CUDA.  (1) register radix-R DIF with bit-reversed outputs, (2) in-place multi-step
DIF in shared memory with per-step twiddles W_{P_{i-1}}^{k_i m}, (3) the two-pass
(four-step) split M = M1 x M2 with inter-pass twiddle W_M^{n2 k1}, (4) the internal
spectrum order S'[k1*M2 + k2] <-> k = k1 + M1*k2.
"""
import numpy as np

PLANS = {2: (2,), 4: (4,), 8: (8,), 16: (16,), 32: (2, 16), 64: (4, 16), 128: (8, 16), 256: (16, 16),
         512: (2, 16, 16), 1024: (4, 16, 16), 2048: (8, 16, 16), 4096: (16, 16, 16)}


def bitrev(k, R):
    b = R.bit_length() - 1
    return int(format(k, "0%db" % b)[::-1], 2) if b else 0


def reg_dif(v):
    """radix-R DIF in 'registers': natural in, X[k] ends in slot bitrev(k)."""
    v = np.array(v, dtype=complex)
    R = len(v)
    ln = R
    while ln >= 2:
        half = ln // 2
        for g in range(0, R, ln):
            for j in range(half):
                a, b = v[g + j], v[g + j + half]
                v[g + j] = a + b
                v[g + j + half] = (a - b) * np.exp(-2j * np.pi * j / ln)
        ln //= 2
    return v


def powers(w, R):
    p = [1.0 + 0j, w] + [0] * (R - 2)
    for k in range(2, R):
        p[k] = p[k // 2] * p[k // 2] if k % 2 == 0 else p[k - 1] * w
    return p[:R]


def smem_fft(x, plan):
    """In-place multi-step DIF exactly as the kernel does it; returns X in natural order
    gathered through the final index formula."""
    L = len(x)
    s = np.array(x, dtype=complex)
    P_prev = L
    out = np.zeros(L, dtype=complex)
    nsteps = len(plan)
    prodR = 1
    for i, R in enumerate(plan):
        P = P_prev // R
        new = s.copy()
        for b in range(L // R):
            h, m = divmod(b, P)
            pos = [h * P_prev + q * P + m for q in range(R)]
            v = reg_dif(s[pos])
            w = np.exp(-2j * np.pi * m / P_prev)
            pw = powers(w, R)
            for k in range(R):
                val = v[bitrev(k, R)]
                if i < nsteps - 1:
                    new[h * P_prev + k * P + m] = val * pw[k]
                else:
                    # h encodes (k_1..k_{s-1}) mixed radix, k_1 most significant
                    hh, kk, mul = h, 0, 1
                    digs = []
                    for Rj in reversed(plan[:-1]):
                        digs.append(hh % Rj); hh //= Rj
                    digs = digs[::-1]  # k_1, k_2, ...
                    for d, Rj in zip(digs, plan[:-1]):
                        kk += d * mul; mul *= Rj
                    out[kk + mul * k] = val
        s = new
        P_prev = P
    return out


def two_pass(x, M1, M2):
    """K1: column FFTs over n1 (stride M2) + twiddle W_M^{n2 k1} -> T[k1][n2];
    K2: row FFTs over n2 -> X[k1 + M1 k2], stored internally at S'[k1*M2+k2]."""
    M = M1 * M2
    x = np.asarray(x, dtype=complex)
    T = np.zeros((M1, M2), dtype=complex)
    for n2 in range(M2):
        col = x[n2::M2]
        Y = smem_fft(col, PLANS[M1])
        T[:, n2] = Y * np.exp(-2j * np.pi * n2 * np.arange(M1) / M)
    Sp = np.zeros(M, dtype=complex)
    for k1 in range(M1):
        Sp[k1 * M2:(k1 + 1) * M2] = smem_fft(T[k1], PLANS[M2])
    X = np.zeros(M, dtype=complex)
    k1, k2 = np.divmod(np.arange(M), M2)
    X[k1 + M1 * k2] = Sp
    return X, Sp


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for L, plan in PLANS.items():
        if L > 1024:
            continue
        x = rng.normal(size=L) + 1j * rng.normal(size=L)
        err = np.max(np.abs(smem_fft(x, plan) - np.fft.fft(x)))
        print("L=%5d plan=%s err=%.2e" % (L, plan, err))
        assert err < 1e-10
    for M1, M2 in ((2, 32), (8, 64), (32, 32), (64, 16)):
        x = rng.normal(size=M1 * M2) + 1j * rng.normal(size=M1 * M2)
        X, Sp = two_pass(x, M1, M2)
        err = np.max(np.abs(X - np.fft.fft(x)))
        print("two-pass %dx%d err=%.2e" % (M1, M2, err))
        assert err < 1e-10
        # final stage: ACF = Re(FFT(S))[n]/M with S read through the internal permutation
        S = np.abs(Sp) ** 2
        n = np.arange(M1 * M2)
        Snat = S[(n % M1) * M2 + n // M1]
        assert np.allclose(Snat, np.abs(np.fft.fft(x)) ** 2)
    print("ok")
