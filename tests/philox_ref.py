"""NumPy restatement of the device's Philox4x32-10 stream layout and fp64 Box-Muller (test only):
counter = (index, t, eval_id lo, eval_id hi ^ stream << 28), key = seed (pf_device.cuh)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed, eval_id, stream, t, index):
    index = np.asarray(index, dtype=np.uint64)
    c0 = index & MASK
    c1 = np.full_like(c0, t)
    c2 = np.full_like(c0, eval_id & 0xFFFFFFFF)
    c3 = np.full_like(c0, ((eval_id >> 32) ^ (stream << 28)) & 0xFFFFFFFF)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normals(seed, eval_id, t, n):
    pairs = np.arange((n + 1) // 2)
    r0, r1, r2, r3 = philox4x32_10(seed, eval_id, 0, t, pairs)
    u1 = ((((r0 << np.uint64(32)) | r1) >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = (((r2 << np.uint64(32)) | r3) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * len(pairs))
    z[0::2] = rad * np.cos(2 * np.pi * u2)
    z[1::2] = rad * np.sin(2 * np.pi * u2)
    return z[:n]


def uniform(seed, eval_id, stream, t, index=0):
    r0, r1, _, _ = philox4x32_10(seed, eval_id, stream, t, np.array([index]))
    return float((((r0 << np.uint64(32)) | r1) >> np.uint64(11))[0]) / 9007199254740992.0
