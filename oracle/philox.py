"""numpy restatement of the library's counter-based Gaussian generator
(csrc/its_common.cuh: philox4x32_10 + philox_normal4).  TEST INFRASTRUCTURE ONLY.

Philox4x32-10 is the published algorithm of Salmon, Moraes, Dror & Shaw,
"Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11); the known-answer vectors
in tests/test_oracle_vs_golden.py are the ones shipped with Random123
(kat_vectors: philox4x32 10 rounds).  The reference itself uses torch's global
generator (torch.randn / randn_like, search_algorithm.py:67, Diffusion.py:96);
the in-kernel stream is a documented replacement for throughput runs, parity runs
inject the reference's tensors.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [..., 4] uint32, key [2] uint32 -> [..., 4] uint32."""
    c = ctr.astype(np.uint64)
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c[..., 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[..., 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def normal(seed: int, cand_id0: int, tag: int, n_units: int, n_per: int) -> np.ndarray:
    """[n_units, n_per] float32 N(0,1): unit i uses counter (quad, cand lo, tag, cand hi)."""
    assert n_per % 4 == 0
    quads = n_per // 4
    ctr = np.zeros((n_units, quads, 4), dtype=np.uint32)
    ctr[..., 0] = np.arange(quads, dtype=np.uint32)[None, :]
    cand = (np.arange(n_units, dtype=np.uint64) + np.uint64(cand_id0))
    ctr[..., 1] = (cand & MASK).astype(np.uint32)[:, None]
    ctr[..., 2] = np.uint32(tag)
    ctr[..., 3] = (cand >> np.uint64(32)).astype(np.uint32)[:, None]
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    inv32 = np.float32(2.3283064365386963e-10)
    out = np.empty((n_units, quads, 4), dtype=np.float32)
    for i in range(2):
        u1 = (r[..., 2 * i].astype(np.float32) + np.float32(1.0)) * inv32
        u2 = r[..., 2 * i + 1].astype(np.float32) * inv32
        rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
        ang = np.float32(6.283185307179586) * u2
        out[..., 2 * i] = rad * np.cos(ang)
        out[..., 2 * i + 1] = rad * np.sin(ang)
    return out.reshape(n_units, n_per)
