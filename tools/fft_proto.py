"""numpy prototype of the device FFT/power algorithm (half-warp per frame).
Validates the index algebra used by csrc/af_fused.cu before it is transcribed to CUDA.
Not part of the product or of the oracle."""
import numpy as np

W = lambda N, m: np.exp(-2j * np.pi * m / N)

def fft4(u0, u1, u2, u3):
    t0, t1, t2, t3 = u0 + u2, u0 - u2, u1 + u3, u1 - u3
    return t0 + t2, t1 - 1j * t3, t0 - t2, t1 + 1j * t3   # k = 0,1,2,3

def fft16(x):
    """x: list of 16 complex (natural order) -> list of 16 (natural order). 4x4 decomposition:
    n = 4a + b, k = c + 4d."""
    Y = [[None] * 4 for _ in range(4)]            # Y[b][c]
    for b in range(4):
        Y[b] = list(fft4(x[b], x[4 + b], x[8 + b], x[12 + b]))
    for b in range(4):
        for c in range(4):
            Y[b][c] = Y[b][c] * W(16, b * c)
    out = [None] * 16
    for c in range(4):
        o = fft4(Y[0][c], Y[1][c], Y[2][c], Y[3][c])
        for d in range(4):
            out[c + 4 * d] = o[d]
    return out

def power_halfwarp(xw):
    """xw: 512 real (windowed, zero padded).  returns P[257] = |rfft|^2 via the lane algorithm."""
    z = xw[0::2] + 1j * xw[1::2]                  # 256 complex
    # pass 1: lane l holds z[16 n1 + l]
    B = np.zeros((16, 16), complex)               # B[lane l][k1]
    for l in range(16):
        A = fft16([z[16 * n1 + l] for n1 in range(16)])
        for k1 in range(16):
            B[l][k1] = A[k1] * W(256, l * k1)
    # transpose through smem: S[k1][n2]
    S = B.T.copy()
    Z = np.zeros((16, 16), complex)               # Z[lane j][k2]  == Zfull[j + 16 k2]
    for j in range(16):
        Z[j] = fft16(list(S[j]))
    Zfull = np.array([Z[k % 16][k // 16] for k in range(256)])
    assert np.allclose(Zfull, np.fft.fft(z)), "256-pt FFT decomposition wrong"
    P = np.full(257, np.nan)
    for j in range(16):
        src = (16 - j) % 16
        pairs = list(range(8)) + ([8] if j == 0 else [])
        for r in pairs:
            zk = Z[j][r]
            if j == 0:
                zp = Z[0][(16 - r) % 16]           # own register (16 - r) & 15
            else:
                zp = Z[src][15 - r]                # shuffle from lane 16 - j, register 15 - r
            k = j + 16 * r
            E2 = zk + np.conj(zp)                  # 2E
            O2 = -1j * (zk - np.conj(zp))          # 2O
            T = W(512, k) * O2
            Sq = (E2.real ** 2 + E2.imag ** 2 + O2.real ** 2 + O2.imag ** 2)
            Cx = 2 * (E2.real * T.real + E2.imag * T.imag)
            P[k] = 0.25 * (Sq + Cx)
            P[256 - k] = 0.25 * (Sq - Cx)
    return P

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = np.zeros(512); x[:400] = rng.standard_normal(400)
    t = [complex(a, b) for a, b in rng.standard_normal((16, 2))]
    assert np.allclose(fft16(t), np.fft.fft(t))
    P = power_halfwarp(x)
    ref = np.abs(np.fft.rfft(x)) ** 2
    assert not np.isnan(P).any()
    print("max rel err", np.max(np.abs(P - ref) / ref.max()))
    assert np.allclose(P, ref)
    print("ok")
