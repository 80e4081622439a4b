"""`two_pass_lanczos_b200.stdrng`: the ChaCha core against published keystreams (zero key, zero nonce: 20 rounds = the
classic ChaCha20 vector, 12 rounds = draft-strombergson-chacha-test-vectors TC1), plus the structural properties of the
seeded uniform stream.  The PCG32 seed expansion and the float conversion are pinned end to end by the reference's published
accuracy rows (tests/test_oracle_reference_kats.py::test_published_accuracy_rows)."""
import numpy as np

from two_pass_lanczos_b200 import stdrng


def _keystream_hex(rounds, nbytes=64, first_block=0):
    words = stdrng.chacha_blocks(np.zeros(8, dtype=np.uint32), first_block, 1, rounds=rounds)[0]
    return words.astype("<u4").tobytes()[:nbytes].hex()


def test_chacha20_zero_key_keystream():
    assert _keystream_hex(20, 32) == "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
    assert _keystream_hex(20, 16, first_block=1) == "9f07e7be5551387a98ba977c732d080d"


def test_chacha12_zero_key_keystream():
    assert _keystream_hex(12, 32) == "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"


def test_blocks_are_counter_addressed_and_vectorised_consistently():
    key = stdrng.seed_from_u64(42)
    many = stdrng.chacha_blocks(key, 0, 7)
    for i in (0, 3, 6):
        assert np.array_equal(many[i], stdrng.chacha_blocks(key, i, 1)[0])
    hi = stdrng.chacha_blocks(key, (1 << 32) - 1, 2)  # the block counter is 64 bits wide
    assert not np.array_equal(hi[0], hi[1]) and not np.array_equal(hi[1], many[0])


def test_seeded_uniform_stream_properties():
    b = stdrng.std_rng_uniform(42, 10_000)
    assert b.shape == (10_000,) and b.min() >= 0.0 and b.max() < 1.0
    assert np.array_equal(b, stdrng.std_rng_uniform(42, 10_000))         # deterministic
    assert np.array_equal(b[:100], stdrng.std_rng_uniform(42, 100))      # a prefix of the same stream
    assert not np.array_equal(b[:100], stdrng.std_rng_uniform(43, 100))
    assert abs(b.mean() - 0.5) < 0.01 and abs(b.var() - 1.0 / 12.0) < 0.005
    assert np.all(b * (1 << 53) == np.floor(b * (1 << 53)))              # 53-bit grid
    assert len(set(stdrng.seed_from_u64(0).tolist())) == 8               # the PCG stream advances before each word
