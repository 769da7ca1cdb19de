"""Oracle (test infrastructure, never imported by the product): NumPy restatement of the counter-based generator the
sampling kernels use for NULL variate pointers (pymc_experimental_b200/csrc/bkf_sample.cu::philox_normal).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11; multipliers
0xD2511F53 / 0xCD9E8D57, Weyl constants 0x9E3779B9 / 0xBB67AE85), checked against the known-answer vectors of the
authors' Random123 library (tests/test_oracle.py).  Variate i of variate
array `stream`: counter = [lo, hi of offset + i // 2, stream, 0], key = [lo, hi of seed]; outputs (r0, r1) and (r2, r3)
give two uniforms ((r_hi << 32 | r_lo) >> 11 + 1/2) 2^-53 in (0, 1); Box-Muller: z_even = sqrt(-2 ln u1) cos(2 pi u2),
z_odd = sqrt(-2 ln u1) sin(2 pi u2).

The reference draws through PyTensor's RandomVariable Ops with a NumPy Generator (filters/distributions.py:52-87); a
draw has no reference numbers to match -- parity of the sampling path is distributional.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32-valued ints, key: (2,) -> (..., 4) outputs (arrays of Python-int-safe uint64)."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return np.stack(c, axis=-1)


def standard_normal(seed, offset, stream, count):
    """Variates 0 .. count-1 of variate array `stream` for (seed, offset), float64."""
    i = np.arange(count, dtype=np.uint64)
    ctr = (np.uint64(offset) + (i >> np.uint64(1)))
    counter = np.stack([ctr & np.uint64(MASK), ctr >> np.uint64(32), np.full_like(ctr, stream), np.zeros_like(ctr)], axis=-1)
    out = philox4x32_10(counter, (seed & MASK, (seed >> 32) & MASK))
    u1 = (((out[..., 0] << np.uint64(32)) | out[..., 1]) >> np.uint64(11)).astype(np.float64) + 0.5
    u2 = (((out[..., 2] << np.uint64(32)) | out[..., 3]) >> np.uint64(11)).astype(np.float64) + 0.5
    u1 *= 2.0 ** -53
    u2 *= 2.0 ** -53
    rad = np.sqrt(-2.0 * np.log(u1))
    return np.where((i & np.uint64(1)) == 1, rad * np.sin(2.0 * np.pi * u2), rad * np.cos(2.0 * np.pi * u2))
