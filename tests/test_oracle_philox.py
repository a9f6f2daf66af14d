"""Pins oracle/philox.py on the Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the Random123
distribution: the all-zero, all-ones and pi-digit cases)."""
import numpy as np

from oracle import philox


def _one(ctr, key):
    r = philox.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
    return tuple(int(x[0]) for x in r)


def test_random123_known_answers():
    assert _one((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert _one((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert _one((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_uniform_range_and_layout():
    u = philox.uniform(1001, seed=1234, scale=2.0, inst=3, upd=7)
    assert u.shape == (1001,) and (u >= 0).all() and (u < 2.0).all()
    assert abs(u.mean() - 1.0) < 0.1
    # odd length is a prefix of the even length; instances / updates give different streams
    assert np.array_equal(u, philox.uniform(1002, 1234, 2.0, 3, 7)[:1001])
    assert not np.array_equal(u, philox.uniform(1001, 1234, 2.0, 4, 7))
    assert not np.array_equal(u, philox.uniform(1001, 1234, 2.0, 3, 8))
    M = philox.uniform_matrix(5, 7, seed=9)
    assert np.array_equal(M.T.ravel(), philox.uniform(35, 9))
