"""CityHash128 v1.0.2 (the ClickHouse frame checksum): the C++ oracle (`oracle/hashes.hpp`) against a second,
structurally different implementation (`tests/cityhash_independent.py`) over every length class of the algorithm
(0, 1-3, 4-8, 9-16, 17-127 with its 16-byte murmur steps, >= 128 with 0-4 tail chunks) and random lengths up
to 64 KiB. The device kernel is checked against both in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cityhash_independent import cityhash128, _len0to16, K2


def test_structural_anchors():
    assert _len0to16(b"") == K2
    assert cityhash128(b"") != cityhash128(b"\x00")


def test_oracle_equals_independent_every_length_class(po):
    rng = np.random.default_rng(20260923)
    lengths = list(range(0, 700)) + [1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097, 30720 + 9, 32768 + 9, 65535]
    lengths += [int(x) for x in rng.integers(700, 40000, size=60)]
    for n in lengths:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert po.cityhash128(data) == cityhash128(data), n
    # low-entropy inputs (runs, which is what compressed frames of constant columns look like)
    for n in (16, 17, 127, 128, 129, 255, 256, 1000, 5000):
        for fill in (b"\x00", b"\xff", b"ab"):
            data = (fill * n)[:n]
            assert po.cityhash128(data) == cityhash128(data), (n, fill)
