"""NumPy restatement of the device-side Brownian increment generator (TEST INFRASTRUCTURE ONLY).

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
known-answer vectors pin it in tests/) keyed by the 64-bit seed, counter = (trajectory lo, trajectory hi,
step, component group), then Box-Muller on (x0, x1) and (x2, x3).  The uint32 stream is exact; the normals go
through NumPy's float32 log / sincos and so agree with the CUDA ones to a few ulp, not bit for bit."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def brownian_increments(seed, t_span, B, D, offset=0):
    """-> dW [T-1, B, D] float32 (see paddlexde_b200/utils/brownian.py)."""
    t = np.asarray(t_span, dtype=np.float32)
    G4 = (D + 3) // 4
    n, b, g = np.meshgrid(np.arange(t.size - 1), np.arange(B) + offset, np.arange(G4), indexing="ij")
    b = b.astype(np.uint64)
    x = philox4x32_10((b & MASK), (b >> np.uint64(32)), n, g, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    k = np.float32(2.0 ** -32)
    f = [xi.astype(np.float32) for xi in x]
    r0 = np.sqrt(np.float32(-2) * np.log(f[0] * k + k, dtype=np.float32), dtype=np.float32)
    r1 = np.sqrt(np.float32(-2) * np.log(f[2] * k + k, dtype=np.float32), dtype=np.float32)
    a0 = (np.float32(2) * (f[1] * k)).astype(np.float64) * np.pi
    a1 = (np.float32(2) * (f[3] * k)).astype(np.float64) * np.pi
    z = np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1)  # [T-1,B,G4,4]
    sq = np.sqrt(np.abs(np.diff(t))).astype(np.float32)
    return (z.reshape(t.size - 1, B, G4 * 4)[..., :D] * sq[:, None, None]).astype(np.float32)
