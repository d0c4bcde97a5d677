"""A second, structurally different CityHash128 (v1.0.2) used ONLY to cross-check `oracle/hashes.hpp` and the
device checksum kernel: written from the published algorithm description (city.cc of CityHash 1.0.2 — the
version ClickHouse froze for its compressed-frame checksums, go-faster/city `CH128`) with Python integers and
byte-slice fetches, no code shared with the C++ oracle or the CUDA kernel.

Known anchors it satisfies by construction of the algorithm: HashLen0to16("") == k2, and for inputs of 16 bytes
or more CityHash128(s) == CityHash128WithSeed(s[16:], (Fetch64(s) ^ k3, Fetch64(s + 8))).
"""
from __future__ import annotations

M64 = (1 << 64) - 1
K0, K1, K2, K3 = 0xc3a5c85c97cb3127, 0xb492b66fbe98f273, 0x9ae16a3b2f90404f, 0xc949d7c7509e6557
KMUL = 0x9ddfea08eb382d69


def _u64(b: bytes, i: int) -> int:
    return int.from_bytes(b[i:i + 8], "little")


def _u32(b: bytes, i: int) -> int:
    return int.from_bytes(b[i:i + 4], "little")


def _ror(v: int, s: int) -> int:
    s &= 63
    return v if s == 0 else ((v >> s) | (v << (64 - s))) & M64


def _mix(v: int) -> int:
    return v ^ (v >> 47)


def _h128to64(lo: int, hi: int) -> int:
    a = ((lo ^ hi) * KMUL) & M64
    a ^= a >> 47
    b = ((hi ^ a) * KMUL) & M64
    b ^= b >> 47
    return (b * KMUL) & M64


def _len0to16(s: bytes) -> int:
    n = len(s)
    if n > 8:
        a, b = _u64(s, 0), _u64(s, n - 8)
        return _h128to64(a, _ror((b + n) & M64, n)) ^ b
    if n >= 4:
        return _h128to64((n + (_u32(s, 0) << 3)) & M64, _u32(s, n - 4))
    if n > 0:
        y = (s[0] + (s[n >> 1] << 8)) & 0xffffffff
        z = (n + (s[n - 1] << 2)) & 0xffffffff
        return (_mix(((y * K2) & M64) ^ ((z * K3) & M64)) * K2) & M64
    return K2


def _weak32(w: int, x: int, y: int, z: int, a: int, b: int):
    a = (a + w) & M64
    b = _ror((b + a + z) & M64, 21)
    c = a
    a = (a + x + y) & M64
    b = (b + _ror(a, 44)) & M64
    return (a + z) & M64, (b + c) & M64


def _weak32_at(s: bytes, i: int, a: int, b: int):
    return _weak32(_u64(s, i), _u64(s, i + 8), _u64(s, i + 16), _u64(s, i + 24), a, b)


def _murmur(s: bytes, seed):
    a, b = seed
    n = len(s)
    if n <= 16:
        a = (_mix((a * K1) & M64) * K1) & M64
        c = ((b * K1) + _len0to16(s)) & M64
        d = _mix((a + (_u64(s, 0) if n >= 8 else c)) & M64)
    else:
        c = _h128to64((_u64(s, n - 8) + K1) & M64, a)
        d = _h128to64((b + n) & M64, (c + _u64(s, n - 16)) & M64)
        a = (a + d) & M64
        i, left = 0, n - 16
        while left > 0:
            a ^= (_mix((_u64(s, i) * K1) & M64) * K1) & M64
            a = (a * K1) & M64
            b ^= a
            c ^= (_mix((_u64(s, i + 8) * K1) & M64) * K1) & M64
            c = (c * K1) & M64
            d ^= c
            i += 16
            left -= 16
    a = _h128to64(a, c)
    b = _h128to64(d, b)
    return a ^ b, _h128to64(b, a)


def cityhash128_with_seed(s: bytes, seed):
    n = len(s)
    if n < 128:
        return _murmur(s, seed)
    x, y = seed
    z = (n * K1) & M64
    v0 = (_ror(y ^ K1, 49) * K1 + _u64(s, 0)) & M64
    v1 = (_ror(v0, 42) * K1 + _u64(s, 8)) & M64
    w0 = (_ror((y + z) & M64, 35) * K1 + x) & M64
    w1 = (_ror((x + _u64(s, 88)) & M64, 53) * K1) & M64
    pos, left = 0, n
    while True:
        for _ in range(2):
            x = (_ror((x + y + v0 + _u64(s, pos + 16)) & M64, 37) * K1) & M64
            y = (_ror((y + v1 + _u64(s, pos + 48)) & M64, 42) * K1) & M64
            x ^= w1
            y ^= v0
            z = _ror(z ^ w0, 33)
            v0, v1 = _weak32_at(s, pos, (v1 * K1) & M64, (x + w0) & M64)
            w0, w1 = _weak32_at(s, pos + 32, (z + w1) & M64, y)
            z, x = x, z
            pos += 64
        left -= 128
        if left < 128:
            break
    y = (y + _ror(w0, 37) * K0 + z) & M64
    x = (x + _ror((v0 + z) & M64, 49) * K0) & M64
    tail = 0
    while tail < left:
        tail += 32
        y = (_ror((y - x) & M64, 42) * K0 + v1) & M64
        w0 = (w0 + _u64(s, pos + left - tail + 16)) & M64
        x = (_ror(x, 49) * K0 + w0) & M64
        w0 = (w0 + v0) & M64
        v0, v1 = _weak32_at(s, pos + left - tail, v0, v1)
    x = _h128to64(x, v0)
    y = _h128to64(y, w0)
    return (_h128to64((x + v1) & M64, w1) + y) & M64, _h128to64((x + w1) & M64, (y + v1) & M64)


def cityhash128(s: bytes):
    """(low64, high64) as ClickHouse stores them in front of a compressed frame (little-endian low, then high)."""
    n = len(s)
    if n >= 16:
        return cityhash128_with_seed(s[16:], (_u64(s, 0) ^ K3, _u64(s, 8)))
    if n >= 8:
        return cityhash128_with_seed(b"", (_u64(s, 0) ^ ((n * K0) & M64), _u64(s, n - 8) ^ K1))
    return cityhash128_with_seed(s, (K0, K1))
