"""Philox oracle against the Random123 known-answer vectors (kat_vectors: philox4x32 10 rounds)."""
import numpy as np

from oracle.philox import philox4x32_10, philox_uniform


def test_random123_kat():
    # Random123 examples/kat_vectors, "philox4x32 10"
    out = philox4x32_10(np.array([0, 0, 0, 0], dtype=np.uint32), np.array([0, 0], dtype=np.uint32))
    assert [hex(int(v)) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = philox4x32_10(np.array([0xFFFFFFFF] * 4, dtype=np.uint32), np.array([0xFFFFFFFF] * 2, dtype=np.uint32))
    assert [hex(int(v)) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    out = philox4x32_10(np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], dtype=np.uint32),
                        np.array([0xA4093822, 0x299F31D0], dtype=np.uint32))
    assert [hex(int(v)) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_uniform_range_and_layout():
    u = philox_uniform(7, 3, 64, 5)
    assert u.shape == (64, 5) and (u >= 0).all() and (u < 1).all()
    # column c of sample i only depends on (i, c): a wider draw has the same leading columns
    u2 = philox_uniform(7, 3, 64, 9)
    assert np.array_equal(u, u2[:, :5])
    assert abs(philox_uniform(1, 0, 4096, 8).mean() - 0.5) < 0.01
