// ORACLE — test infrastructure only. Never linked into or called from the product path.
//
// SHA-256 / HMAC-SHA256 (FIPS 180-4, RFC 2104) as used by
//   pkg/transformer/registry/mask/hmac_hasher.go:29-33  (hmac.New(sha256.New, salt); hex)
// and CityHash128 v1.0.2, the checksum in front of every ClickHouse compressed frame
// (third-party: github.com/go-faster/city v1.0.1 `CH128`, reached through
//  github.com/ClickHouse/ch-go v0.71.0 compress.Writer; reference call sites
//  pkg/providers/clickhouse/conn/connection.go:46, sink_table.go:657-674).
// PARITY UNPINNED for CityHash128: the reference holds no known-answer vector for it
// (sink tests use sqlmock.AnyArg()); this follows the published CityHash 1.0.2 source.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>

namespace orc {

// ------------------------------------------------------------------ SHA-256
struct Sha256 {
    uint32_t h[8]; uint8_t buf[64]; uint64_t len; int fill;
    static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void init() {
        static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        std::memcpy(h, iv, sizeof iv); len = 0; fill = 0;
    }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
            0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
            0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
            0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + K[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const uint8_t* p, size_t n) {
        len += n;
        while (n) {
            size_t k = 64 - fill; if (k > n) k = n;
            std::memcpy(buf + fill, p, k); fill += (int)k; p += k; n -= k;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad = 0x80; update(&pad, 1);
        uint8_t z = 0; while (fill != 56) update(&z, 1);
        uint8_t lb[8]; for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
    }
};

// crypto/hmac: key longer than the block is hashed first; fresh pads per value, as the
// reference does (hmac.New per value, hmac_hasher.go:30).
inline void hmac_sha256(const uint8_t* key, size_t klen, const uint8_t* msg, size_t mlen, uint8_t out[32]) {
    uint8_t k[64]; std::memset(k, 0, 64);
    if (klen > 64) { Sha256 s; s.init(); s.update(key, klen); s.final(k); }
    else std::memcpy(k, key, klen);
    uint8_t ipad[64], opad[64];
    for (int i = 0; i < 64; i++) { ipad[i] = k[i] ^ 0x36; opad[i] = k[i] ^ 0x5c; }
    uint8_t inner[32];
    Sha256 s; s.init(); s.update(ipad, 64); s.update(msg, mlen); s.final(inner);
    s.init(); s.update(opad, 64); s.update(inner, 32); s.final(out);
}
inline std::string hex_lower(const uint8_t* p, size_t n) {
    static const char* H = "0123456789abcdef"; std::string s; s.resize(2 * n);
    for (size_t i = 0; i < n; i++) { s[2 * i] = H[p[i] >> 4]; s[2 * i + 1] = H[p[i] & 15]; }
    return s;
}

// ------------------------------------------------------------------ CityHash128 v1.0.2
namespace city {
typedef std::pair<uint64_t, uint64_t> u128;   // (low, high)
static const uint64_t k0 = 0xc3a5c85c97cb3127ULL, k1 = 0xb492b66fbe98f273ULL, k2 = 0x9ae16a3b2f90404fULL, k3 = 0xc949d7c7509e6557ULL;
inline uint64_t f64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint32_t f32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
inline uint64_t rot1(uint64_t v, int s) { return (v >> s) | (v << (64 - s)); }
inline uint64_t smix(uint64_t v) { return v ^ (v >> 47); }
inline uint64_t h128to64(uint64_t lo, uint64_t hi) {
    const uint64_t kMul = 0x9ddfea08eb382d69ULL;
    uint64_t a = (lo ^ hi) * kMul; a ^= (a >> 47);
    uint64_t b = (hi ^ a) * kMul; b ^= (b >> 47); b *= kMul; return b;
}
inline uint64_t hl16(uint64_t u, uint64_t v) { return h128to64(u, v); }
inline uint64_t hl0to16(const uint8_t* s, size_t len) {
    if (len > 8) { uint64_t a = f64(s), b = f64(s + len - 8); return hl16(a, rot1(b + len, (int)len)) ^ b; }
    if (len >= 4) { uint64_t a = f32(s); return hl16(len + (a << 3), f32(s + len - 4)); }
    if (len > 0) {
        uint8_t a = s[0], b = s[len >> 1], c = s[len - 1];
        uint32_t y = (uint32_t)a + ((uint32_t)b << 8); uint32_t z = (uint32_t)len + ((uint32_t)c << 2);
        return smix(y * k2 ^ z * k3) * k2;
    }
    return k2;
}
inline u128 weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a, uint64_t b) {
    a += w; b = rot(b + a + z, 21); uint64_t c = a; a += x; a += y; b += rot(a, 44); return u128(a + z, b + c);
}
inline u128 weak32(const uint8_t* s, uint64_t a, uint64_t b) { return weak32(f64(s), f64(s + 8), f64(s + 16), f64(s + 24), a, b); }
inline u128 murmur(const uint8_t* s, size_t len, u128 seed) {
    uint64_t a = seed.first, b = seed.second, c = 0, d = 0;
    long l = (long)len - 16;
    if (l <= 0) {
        a = smix(a * k1) * k1; c = b * k1 + hl0to16(s, len); d = smix(a + (len >= 8 ? f64(s) : c));
    } else {
        c = hl16(f64(s + len - 8) + k1, a); d = hl16(b + len, c + f64(s + len - 16)); a += d;
        do { a ^= smix(f64(s) * k1) * k1; a *= k1; b ^= a; c ^= smix(f64(s + 8) * k1) * k1; c *= k1; d ^= c; s += 16; l -= 16; } while (l > 0);
    }
    a = hl16(a, c); b = hl16(d, b);
    return u128(a ^ b, hl16(b, a));
}
inline u128 hash128_seed(const uint8_t* s, size_t len, u128 seed) {
    if (len < 128) return murmur(s, len, seed);
    u128 v, w; uint64_t x = seed.first, y = seed.second, z = len * k1;
    v.first = rot(y ^ k1, 49) * k1 + f64(s);
    v.second = rot(v.first, 42) * k1 + f64(s + 8);
    w.first = rot(y + z, 35) * k1 + x;
    w.second = rot(x + f64(s + 88), 53) * k1;
    do {
        for (int rep = 0; rep < 2; rep++) {
            x = rot(x + y + v.first + f64(s + 16), 37) * k1;
            y = rot(y + v.second + f64(s + 48), 42) * k1;
            x ^= w.second; y ^= v.first; z = rot(z ^ w.first, 33);
            v = weak32(s, v.second * k1, x + w.first);
            w = weak32(s + 32, z + w.second, y);
            std::swap(z, x); s += 64;
        }
        len -= 128;
    } while (len >= 128);
    y += rot(w.first, 37) * k0 + z;
    x += rot(v.first + z, 49) * k0;
    for (size_t tail = 0; tail < len;) {
        tail += 32;
        y = rot(y - x, 42) * k0 + v.second;
        w.first += f64(s + len - tail + 16);
        x = rot(x, 49) * k0 + w.first;
        w.first += v.first;
        v = weak32(s + len - tail, v.first, v.second);
    }
    x = hl16(x, v.first); y = hl16(y, w.first);
    return u128(hl16(x + v.second, w.second) + y, hl16(x + w.second, y + v.second));
}
inline u128 hash128(const uint8_t* s, size_t len) {
    if (len >= 16) return hash128_seed(s + 16, len - 16, u128(f64(s) ^ k3, f64(s + 8)));
    if (len >= 8) return hash128_seed(nullptr, 0, u128(f64(s) ^ (len * k0), f64(s + len - 8) ^ k1));
    return hash128_seed(s, len, u128(k0, k1));
}
}  // namespace city
}  // namespace orc
