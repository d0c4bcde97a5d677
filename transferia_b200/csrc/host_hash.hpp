// Host-side CityHash128 (v1.0.2, the variant ClickHouse keeps for its compressed-frame checksums) and an LZ4 block decoder.
// Used by the ClickHouse wire writer for the few frames the HOST has to produce or read itself: the empty blocks that open and
// close an INSERT, and the compressed sample / data blocks the server sends back. The frames of real data come from the device
// (kernels_lz4.cuh) and are only forwarded.
#pragma once
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace tfh {

using u128 = std::pair<uint64_t, uint64_t>;   // (low, high) as ClickHouse writes them: low first

namespace city {
constexpr uint64_t K0 = 0xc3a5c85c97cb3127ULL, K1 = 0xb492b66fbe98f273ULL, K2 = 0x9ae16a3b2f90404fULL, K3 = 0xc949d7c7509e6557ULL;
inline uint64_t ld64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint32_t ld32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t ror(uint64_t v, int s) { return s == 0 ? v : (v >> s) | (v << (64 - s)); }
inline uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
inline uint64_t len16(uint64_t u, uint64_t v) {
    const uint64_t mul = 0x9ddfea08eb382d69ULL;
    uint64_t a = (u ^ v) * mul; a ^= a >> 47;
    uint64_t b = (v ^ a) * mul; b ^= b >> 47;
    return b * mul;
}
inline uint64_t len0to16(const uint8_t* s, size_t n) {
    if (n > 8) { const uint64_t a = ld64(s), b = ld64(s + n - 8); return len16(a, ror(b + n, (int)n)) ^ b; }
    if (n >= 4) { const uint64_t a = ld32(s); return len16(n + (a << 3), ld32(s + n - 4)); }
    if (n > 0) {
        const uint32_t y = (uint32_t)s[0] + ((uint32_t)s[n >> 1] << 8), z = (uint32_t)n + ((uint32_t)s[n - 1] << 2);
        return shift_mix(y * K2 ^ z * K3) * K2;
    }
    return K2;
}
inline u128 weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a, uint64_t b) {
    a += w; b = ror(b + a + z, 21); const uint64_t c = a; a += x; a += y; b += ror(a, 44);
    return {a + z, b + c};
}
inline u128 weak32(const uint8_t* s, uint64_t a, uint64_t b) { return weak32(ld64(s), ld64(s + 8), ld64(s + 16), ld64(s + 24), a, b); }
inline u128 murmur(const uint8_t* s, size_t n, u128 seed) {
    uint64_t a = seed.first, b = seed.second, c, d;
    long l = (long)n - 16;
    if (l <= 0) {
        a = shift_mix(a * K1) * K1; c = b * K1 + len0to16(s, n); d = shift_mix(a + (n >= 8 ? ld64(s) : c));
    } else {
        c = len16(ld64(s + n - 8) + K1, a); d = len16(b + n, c + ld64(s + n - 16)); a += d;
        do {
            a ^= shift_mix(ld64(s) * K1) * K1; a *= K1; b ^= a;
            c ^= shift_mix(ld64(s + 8) * K1) * K1; c *= K1; d ^= c;
            s += 16; l -= 16;
        } while (l > 0);
    }
    a = len16(a, c); b = len16(d, b);
    return {a ^ b, len16(b, a)};
}
inline u128 with_seed(const uint8_t* s, size_t n, u128 seed) {
    if (n < 128) return murmur(s, n, seed);
    u128 v, w; uint64_t x = seed.first, y = seed.second, z = n * K1;
    v.first = ror(y ^ K1, 49) * K1 + ld64(s);
    v.second = ror(v.first, 42) * K1 + ld64(s + 8);
    w.first = ror(y + z, 35) * K1 + x;
    w.second = ror(x + ld64(s + 88), 53) * K1;
    do {
        for (int half = 0; half < 2; half++) {
            x = ror(x + y + v.first + ld64(s + 16), 37) * K1;
            y = ror(y + v.second + ld64(s + 48), 42) * K1;
            x ^= w.second; y ^= v.first; z = ror(z ^ w.first, 33);
            v = weak32(s, v.second * K1, x + w.first);
            w = weak32(s + 32, z + w.second, y);
            std::swap(z, x); s += 64;
        }
        n -= 128;
    } while (n >= 128);
    y += ror(w.first, 37) * K0 + z;
    x += ror(v.first + z, 49) * K0;
    for (size_t done = 0; done < n;) {
        done += 32;
        y = ror(y - x, 42) * K0 + v.second;
        w.first += ld64(s + n - done + 16);
        x = ror(x, 49) * K0 + w.first;
        w.first += v.first;
        v = weak32(s + n - done, v.first, v.second);
    }
    x = len16(x, v.first); y = len16(y, w.first);
    return {len16(x + v.second, w.second) + y, len16(x + w.second, y + v.second)};
}
}  // namespace city

inline u128 cityhash128(const uint8_t* s, size_t n) {
    using namespace city;
    if (n >= 16) return with_seed(s + 16, n - 16, {ld64(s) ^ K3, ld64(s + 8)});
    if (n >= 8) return with_seed(nullptr, 0, {ld64(s) ^ (n * K0), ld64(s + n - 8) ^ K1});
    return with_seed(s, n, {K0, K1});
}

// LZ4 block format decoder with full bounds checks; false on any malformed input or when the output is not exactly `want` bytes.
inline bool lz4_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t want) {
    size_t ip = 0, op = 0;
    while (ip < n) {
        const uint8_t tok = src[ip++];
        size_t lit = tok >> 4;
        if (lit == 15) { uint8_t b; do { if (ip >= n) return false; b = src[ip++]; lit += b; } while (b == 255); }
        if (lit > n - ip || lit > want - op) return false;
        std::memcpy(dst + op, src + ip, lit); ip += lit; op += lit;
        if (ip == n) break;                                     // the last sequence has no match part
        if (n - ip < 2) return false;
        const size_t off = (size_t)src[ip] | ((size_t)src[ip + 1] << 8); ip += 2;
        if (off == 0 || off > op) return false;
        size_t ml = (size_t)(tok & 15) + 4;
        if ((tok & 15) == 15) { uint8_t b; do { if (ip >= n) return false; b = src[ip++]; ml += b; } while (b == 255); }
        if (ml > want - op) return false;
        for (size_t k = 0; k < ml; k++) dst[op + k] = dst[op + k - off];   // overlapping copies repeat the pattern
        op += ml;
    }
    return op == want;
}

// One ClickHouse compressed frame around `raw` with a literal-only LZ4 body (valid for any input; the host only frames tiny blocks).
inline void frame_literal(const uint8_t* raw, size_t n, std::vector<uint8_t>& out) {
    std::vector<uint8_t> body;
    size_t lit = n;
    body.push_back((uint8_t)((lit >= 15 ? 15 : lit) << 4));
    if (lit >= 15) { lit -= 15; while (lit >= 255) { body.push_back(255); lit -= 255; } body.push_back((uint8_t)lit); }
    body.insert(body.end(), raw, raw + n);
    const uint32_t csize = (uint32_t)(9 + body.size()), usize = (uint32_t)n;
    std::vector<uint8_t> f; f.reserve(csize);
    f.push_back(0x82);
    f.insert(f.end(), (const uint8_t*)&csize, (const uint8_t*)&csize + 4);
    f.insert(f.end(), (const uint8_t*)&usize, (const uint8_t*)&usize + 4);
    f.insert(f.end(), body.begin(), body.end());
    const u128 h = cityhash128(f.data(), f.size());
    const size_t at = out.size(); out.resize(at + 16 + f.size());
    std::memcpy(&out[at], &h.first, 8); std::memcpy(&out[at + 8], &h.second, 8);
    std::memcpy(&out[at + 16], f.data(), f.size());
}

}  // namespace tfh
