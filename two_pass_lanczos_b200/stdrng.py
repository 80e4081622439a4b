"""`StdRng::seed_from_u64(seed)` + `rng.random::<f64>()` of rand 0.9.2 restated: the right-hand sides of
the reference's tests and benches (`tests/correctness.rs:109-110`, `src/algorithms/mod.rs:439`, `src/bin/stability.rs:257`,
`src/bin/orthogonality.rs:163`, `src/bin/dense_tradeoff.rs:156-157`) are `StdRng::seed_from_u64(42)` uniforms on [0, 1).

What this follows (rand 0.9.2 / rand_chacha 0.9.0 / rand_core 0.9.x are registry dependencies, `Cargo.lock:1097-1119`, not in
the tree -- restated from their published algorithms):
  * StdRng = ChaCha12: state = "expand 32-byte k" | key (8 words, the seed) | 64-bit block counter (words 12-13, from 0) |
    64-bit stream id (words 14-15, 0); 6 double rounds; output = state + input; blocks are consumed word by word.
  * `seed_from_u64`: the 32-byte seed is filled 4 bytes at a time from a PCG32 stream (multiplier 6364136223846793005,
    increment 11634580027462260723, state advanced BEFORE each output, output = rotr32(((s >> 18) ^ s) >> 27, s >> 59)).
  * `next_u64` = two consecutive words, low word first; f64 = (next_u64 >> 11) * 2^-53.
PINNING: the ChaCha core is checked against the published 20-round and 12-round zero-key keystreams
(tests/test_stdrng_cpu.py).  Seed expansion, word order and float conversion are pinned END TO END by outputs of the reference
itself: with this b the CPU restatement of the reference path (the checker used by tests/) reproduces every row of the reference's results/accuracy_*.csv (produced by
src/bin/stability.rs with `StdRng::seed_from_u64(42)`) to 1e-14 ... 1e-10 relative on the printed errors, where any other
right-hand side is off by 3-50 % (test_published_accuracy_rows in tests/).
"""
from __future__ import annotations

import numpy as np

_MASK64 = (1 << 64) - 1
_CONSTANTS = np.array([0x61707865, 0x3320646E, 0x79622D32, 0x6B206574], dtype=np.uint32)


def _rotl(x, n):
    return (x << np.uint32(n)) | (x >> np.uint32(32 - n))


def _quarter(s, a, b, c, d):
    s[a] += s[b]; s[d] ^= s[a]; s[d] = _rotl(s[d], 16)
    s[c] += s[d]; s[b] ^= s[c]; s[b] = _rotl(s[b], 12)
    s[a] += s[b]; s[d] ^= s[a]; s[d] = _rotl(s[d], 8)
    s[c] += s[d]; s[b] ^= s[c]; s[b] = _rotl(s[b], 7)


def chacha_blocks(key_words, first_block: int, nblocks: int, rounds: int = 12, stream: int = 0) -> np.ndarray:
    """uint32[nblocks, 16] keystream words of blocks first_block .. first_block + nblocks - 1 (vectorised over the blocks)."""
    key = np.asarray(key_words, dtype=np.uint32)
    if key.shape != (8,) or rounds % 2:
        raise ValueError("key must be 8 words, rounds even")
    ctr = (np.arange(nblocks, dtype=np.uint64) + np.uint64(first_block))
    init = [np.full(nblocks, c, dtype=np.uint32) for c in _CONSTANTS]
    init += [np.full(nblocks, k, dtype=np.uint32) for k in key]
    init += [(ctr & np.uint64(0xFFFFFFFF)).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32)]
    init += [np.full(nblocks, stream & 0xFFFFFFFF, dtype=np.uint32), np.full(nblocks, (stream >> 32) & 0xFFFFFFFF, dtype=np.uint32)]
    s = [w.copy() for w in init]
    with np.errstate(over="ignore"):
        for _ in range(rounds // 2):
            _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13); _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
            _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12); _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
        out = np.stack([a + b for a, b in zip(s, init)], axis=1)
    return out


def seed_from_u64(state: int) -> np.ndarray:
    """The 8 key words `SeedableRng::seed_from_u64` derives from a u64 (PCG32 expansion)."""
    mul, inc = 6364136223846793005, 11634580027462260723
    words = []
    for _ in range(8):
        state = (state * mul + inc) & _MASK64
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF)
    return np.array(words, dtype=np.uint32)  # to_le_bytes + little-endian key words = the values themselves


def std_rng_u32(seed: int, count: int) -> np.ndarray:
    """the first `count` `next_u32` outputs of `StdRng::seed_from_u64(seed)`"""
    nblocks = (count + 15) // 16
    return chacha_blocks(seed_from_u64(seed), 0, max(nblocks, 1), rounds=12).reshape(-1)[:count]


def std_rng_uniform(seed: int, n: int) -> np.ndarray:
    """n draws of `rng.random::<f64>()` from `StdRng::seed_from_u64(seed)`: uniform on [0, 1) with 53 random bits"""
    w = std_rng_u32(seed, 2 * n).astype(np.uint64)
    u64 = w[0::2] | (w[1::2] << np.uint64(32))
    return (u64 >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
