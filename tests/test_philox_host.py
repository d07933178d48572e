"""CPU: the host restatement of Philox4x32 (7 and 10 rounds) (tests/philox_host.py) against the Random123 known-answer vectors."""
import numpy as np

from philox_host import mulhi, philox4x32

KAT = [  # Random123 kat_vectors: philox4x32 10 rounds — counter, key -> output
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32(*ctr, *key)
        assert tuple(int(x) for x in got) == want


KAT7 = [  # Random123 kat_vectors: philox4x32 7 rounds (the per-step slip / noise draws use the 7-round variant)
    ((0, 0, 0, 0), (0, 0), (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)),
]


def test_philox_7_rounds_known_answers():
    for ctr, key, want in KAT7:
        got = philox4x32(*ctr, *key, rounds=7)
        assert tuple(int(x) for x in got) == want


def test_philox_vectorised_matches_scalar():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2**32, size=(4, 257), dtype=np.uint64)
    out = philox4x32(c[0], c[1], c[2], c[3], 0x12345678, 0x9abcdef0)
    for i in (0, 100, 256):
        one = philox4x32(int(c[0, i]), int(c[1, i]), int(c[2, i]), int(c[3, i]), 0x12345678, 0x9abcdef0)
        assert [int(x[i]) for x in out] == [int(x) for x in one]


def test_mulhi_bounds():
    u = np.array([0, 1, 2**31, 2**32 - 1], dtype=np.uint64)
    assert mulhi(u, 300).tolist() == [0, 0, 150, 299]
