// LZ4 block compression of the native block, cut into ClickHouse compressed frames, plus the
// CityHash128 (v1.0.2) frame checksum and the final gather into one contiguous wire buffer.
//
// GPU-native LZ4 (not a port of any CPU compressor): one CTA per frame, the frame lives in shared
// memory, and every phase is data-parallel:
//   P1  stage the frame in shared memory (16-byte coalesced loads)
//   P2  match finding: 2048 positions per round (4 consecutive per thread: two shared-memory words give
//       the four 4-byte sequences) probe a 4096-entry u16 hash table in shared memory, ONE barrier,
//       then insert; a probe therefore only ever sees earlier rounds (or, harmlessly, racing inserts of
//       this round, filtered by `cand < pos`); candidates are verified against the data immediately;
//       result = 1 valid bit + u16 candidate per position
//   P3  greedy parse, one thread per 64-byte segment (matches are cut at the segment end), sequence
//       descriptors overwrite the segment's own candidate slots in place
//   P4  segmented scan carries pending literals across segments; block scan gives every segment its
//       output offset; a suffix-min tells each segment which later sequence owns its trailing literals
//   P5  all threads emit tokens/lengths/offsets and copy their own segment's literals
// The emitted stream is a standard LZ4 block (last 5 bytes literals, last match starts >= 12 bytes
// before the end) and decodes with stock liblz4; the bytes are NOT those of pierrec/lz4 (parity for
// compressed bytes is unpinned in the reference).
#pragma once
#include "device_types.cuh"
#include "kernels_encode.cuh"

namespace tfk {

#define LZ_THREADS 512
#define LZ_SEG 64
#define LZ_HASH_BITS 11      /* 2048 entries of u32: tag16 << 16 | position */
#define LZ_MAX_FRAME 32768
#define LZ_HDR 25          // 16 checksum + 1 method + 4 compressed size + 4 raw size

__host__ __device__ inline uint32_t lz4_bound(uint32_t n) { return n + n / 255 + 16; }
__host__ __device__ inline uint32_t lz_slot_stride(uint32_t frame_bytes) { return (LZ_HDR + 7 + lz4_bound(frame_bytes) + 15) & ~15u; }
// slot layout: [7 pad][16 checksum][0x82][u32][u32][lz4 block]; the checksum field starts at +7 so that
// the hashed region (+23) ... keep it simple: the slot starts 16-byte aligned and the LZ4 block at +25.

struct Lz4Args {
    const uint8_t* raw; DState* st; uint8_t* slots; uint32_t slot_stride; uint32_t* comp_size; uint32_t frame_bytes;
    unsigned long long* phases;     // optional [8]: cycles thread 0 of every CTA spent in P1..P5 (tfgpu_debug_lz4_phases), NULL = off
};
#define LZ_PHASE(k) do { if (a.phases && tid == 0) { const long long t_ = clock64(); atomicAdd(&a.phases[k], (unsigned long long)(t_ - t_ph)); t_ph = t_; } } while (0)

// The frame is kept in shared memory with one padding word after every 16 words (64 bytes): the parser and the emitter give
// every thread its own 64-byte segment, so without the padding the 32 lanes of a warp would hit only two banks.
#define LZ_PW(i) ((i) + ((i) >> 4))                 /* padded word index of data word i */
#define LZ_PB(p) ((p) + (((p) >> 6) << 2))          /* padded byte address of data byte p */
__host__ __device__ inline uint32_t lz_data_bytes(uint32_t frame_bytes) { return (frame_bytes + frame_bytes / 16 + 32 + 15) & ~15u; }
__device__ __forceinline__ uint32_t ld32u(const uint32_t* w, uint32_t p) {   // 4 bytes at byte offset p of the (padded) frame in shared memory
    const uint32_t i = p >> 2, s = (p & 3) * 8;
    return __funnelshift_r(w[LZ_PW(i)], w[LZ_PW(i + 1)], s);
}
__device__ __forceinline__ uint32_t ext_bytes(uint32_t x) { return x < 15 ? 0u : 1u + (x - 15u) / 255u; }
__device__ __forceinline__ uint8_t* put_ext(uint8_t* o, uint32_t x) {   // x >= 15
    x -= 15; while (x >= 255) { *o++ = 255; x -= 255; } *o++ = (uint8_t)x; return o;
}
// literal copy shared -> global: bytes up to a 4-byte boundary of the destination, then aligned words
// (source re-aligned with a funnel shift), then the tail
__device__ __forceinline__ void copy_s2g(uint8_t* dst, const uint32_t* data_w, uint32_t src, uint32_t n) {
    const uint8_t* data = (const uint8_t*)data_w;
    while (n && ((uintptr_t)dst & 3)) { *dst++ = data[LZ_PB(src)]; src++; n--; }
    for (; n >= 4; n -= 4, dst += 4, src += 4) *(uint32_t*)dst = ld32u(data_w, src);
    while (n) { *dst++ = data[LZ_PB(src)]; src++; n--; }
}

#ifdef TF_KERNELS_LZ4
__global__ void __launch_bounds__(LZ_THREADS, 2) k_lz4_frames(Lz4Args a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t F = a.frame_bytes;
    const uint32_t DB = lz_data_bytes(F);
    uint32_t* data_w = (uint32_t*)smem;                                  // the frame, padded (LZ_PW / LZ_PB)
    uint8_t* data = smem;
    uint16_t* cand = (uint16_t*)(smem + DB);                             // 2F bytes; later: sequence descriptors
    uint32_t* table = (uint32_t*)(smem + DB + 2 * F);                   // 8 KB; later: per-segment arrays
    uint8_t* vbits = smem + DB + 2 * F + (4u << LZ_HASH_BITS);          // F/8 bytes: 1 candidate bit per position
    uint32_t* scratch = (uint32_t*)(vbits + F / 8);                      // 64 words
    __shared__ uint32_t s_frame;
    // per-segment arrays aliased onto the hash table after P2 (nseg <= 512)
    uint32_t* seg_off = (uint32_t*)table;            // [513]
    int32_t* delta0 = (int32_t*)(seg_off + 516);     // [512]
    uint16_t* carry_incl = (uint16_t*)(delta0 + 512);   // [512]
    uint16_t* next_has = carry_incl + 512;           // [512]
    uint8_t* nseq_s = (uint8_t*)(next_has + 512);    // [512]
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t raw_total = a.st->raw_total, n_frames = a.st->n_frames;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_frame = atomicAdd(&a.st->frame_ticket, 1u);
        __syncthreads();
        const uint32_t f = s_frame;
        if (f >= n_frames) break;
        long long t_ph = a.phases ? clock64() : 0;
        const uint64_t pos0 = (uint64_t)f * F;
        const uint32_t len = (uint32_t)((raw_total - pos0 < F) ? raw_total - pos0 : F);
        uint8_t* out = a.slots + (size_t)f * a.slot_stride + LZ_HDR;

        // ---- P1: stage
        {
            const int4* g = (const int4*)(a.raw + pos0);
            const uint32_t nv = (len + 15) >> 4;
            for (uint32_t i = tid; i < (F + 16) / 16; i += LZ_THREADS) {
                int4 v = make_int4(0, 0, 0, 0);
                if (i < nv) v = __ldg(g + i);
                uint32_t* d = data_w + LZ_PW(4 * i);                      // the 4 words of a chunk share one 16-word group
                d[0] = (uint32_t)v.x; d[1] = (uint32_t)v.y; d[2] = (uint32_t)v.z; d[3] = (uint32_t)v.w;
            }
            for (uint32_t i = tid; i < (1u << LZ_HASH_BITS); i += LZ_THREADS) table[i] = 0;
        }
        __syncthreads();
        if (len & 15) {   // zero the bytes past len inside the last 16-byte chunk (they belong to the next frame)
            if (tid < 16 && (len & ~15u) + tid >= len) data[LZ_PB((len & ~15u) + tid)] = 0;
        }
        __syncthreads();

        LZ_PHASE(0);
        // ---- P2: match finding
        const uint32_t nrounds = (len + 2047) / 2048;
        for (uint32_t rd = 0; rd < nrounds; rd++) {
            const uint32_t p0 = rd * 2048 + tid * 4;
            const bool active = p0 < F;                         // F is a multiple of 64, so p0 + 3 < F too
            uint32_t idx[4], tag[4], c[4];
            if (active) {
                const uint32_t w0 = data_w[LZ_PW(p0 >> 2)], w1 = data_w[LZ_PW((p0 >> 2) + 1)];
                uint32_t seq[4];
                seq[0] = w0; seq[1] = __funnelshift_r(w0, w1, 8); seq[2] = __funnelshift_r(w0, w1, 16); seq[3] = __funnelshift_r(w0, w1, 24);
#pragma unroll
                for (int k = 0; k < 4; k++) { const uint32_t h = seq[k] * 2654435761u; idx[k] = h >> (32 - LZ_HASH_BITS); tag[k] = (h << LZ_HASH_BITS) & 0xffff0000u; c[k] = table[idx[k]]; }
            }
            __syncthreads();
            uint32_t vm = 0;      // candidate bits: the 27 known hash bits agree; the bytes are compared by the parser (P3)
            if (active) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t p = p0 + k;
                    if (p < len) table[idx[k]] = tag[k] | p;
                    if ((p + 12 <= len) && (c[k] & 0xffff0000u) == tag[k] && (c[k] & 0xffffu) < p) vm |= 1u << k;   // MFLIMIT
                }
            }
            __syncthreads();
            if (active) {   // second probe: sees this round's inserts, recovers repeats whose first occurrence is in this round
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t p = p0 + k;
                    if (!((vm >> k) & 1) && p + 12 <= len) {
                        const uint32_t c2 = table[idx[k]];
                        if ((c2 & 0xffff0000u) == tag[k] && (c2 & 0xffffu) < p) { vm |= 1u << k; c[k] = c2; }
                    }
                }
                *(uint2*)(cand + p0) = make_uint2((c[0] & 0xffffu) | (c[1] << 16), (c[2] & 0xffffu) | (c[3] << 16));
            }
            // two lanes share one byte of the bit map
            const uint32_t other = __shfl_down_sync(0xffffffffu, vm, 1);
            if (active && !(lane & 1)) vbits[p0 >> 3] = (uint8_t)(vm | (other << 4));
        }
        __syncthreads();

        LZ_PHASE(1);
        // ---- P3: greedy parse, one thread per 64-byte segment
        const uint32_t nseg = (len + LZ_SEG - 1) / LZ_SEG;
        uint32_t my_nseq = 0, my_trail = 0, seg_start = tid * LZ_SEG, seg_end = 0;
        uint2* desc = (uint2*)(cand + seg_start);
        if (tid < nseg) {
            seg_end = seg_start + LZ_SEG < len ? seg_start + LZ_SEG : len;
            const uint32_t lim5 = len >= 5 ? len - 5 : 0;
            const uint32_t limit = seg_end < lim5 ? seg_end : lim5;       // matches end before the last 5 bytes and inside the segment
            const uint64_t m64 = *(const uint64_t*)(vbits + 8 * tid);      // bit i = position seg_start + i has a candidate
            uint32_t cur = 0, last_end = seg_start;
            while (cur < LZ_SEG) {
                const uint64_t mm = m64 >> cur;
                if (!mm) break;
                const uint32_t r = cur + (uint32_t)__ffsll((long long)mm) - 1;
                const uint32_t p = seg_start + r;
                if (p + 4 > limit) break;
                const uint32_t c = cand[p];
                const uint32_t maxl = limit - p;
                uint32_t ml = 0;
                while (ml < maxl) {
                    const uint32_t x = ld32u(data_w, c + ml) ^ ld32u(data_w, p + ml);
                    if (x) { ml += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
                    ml += 4;
                }
                if (ml > maxl) ml = maxl;
                if (ml < 4) { cur = r + 1; continue; }            // the tag agreed but the bytes do not: not a match
                desc[my_nseq] = make_uint2(p | (ml << 16), p - c);
                my_nseq++; cur = r + ml; last_end = p + ml;
            }
            my_trail = seg_end - last_end;
        }
        __syncthreads();   // everyone is done reading the hash table region? (P2 finished before P3) -> reuse it now

        LZ_PHASE(2);
        // ---- P4a: segmented scan of pending literals: combine(a, b) = b.has ? b : (a.has, a.tr + b.tr)
        {
            uint32_t has = my_nseq > 0, tr = my_trail;
            if (tid >= nseg) { has = 0; tr = 0; }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t h2 = __shfl_up_sync(0xffffffffu, has, d), t2 = __shfl_up_sync(0xffffffffu, tr, d);
                if (lane >= (uint32_t)d && !has) { tr += t2; has = h2; }
            }
            if (lane == 31) { scratch[warp] = has; scratch[16 + warp] = tr; }
            __syncthreads();
            if (warp == 0) {
                uint32_t wh = lane < LZ_THREADS / 32 ? scratch[lane] : 0, wt = lane < LZ_THREADS / 32 ? scratch[16 + lane] : 0;
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    const uint32_t h2 = __shfl_up_sync(0xffffffffu, wh, d), t2 = __shfl_up_sync(0xffffffffu, wt, d);
                    if (lane >= (uint32_t)d && !wh) { wt += t2; wh = h2; }
                }
                // exclusive prefix for warp w = inclusive of w-1
                const uint32_t eh = __shfl_up_sync(0xffffffffu, wh, 1), et = __shfl_up_sync(0xffffffffu, wt, 1);
                if (lane < LZ_THREADS / 32) { scratch[lane] = lane ? eh : 0; scratch[16 + lane] = lane ? et : 0; }
            }
            __syncthreads();
            if (!has) { tr += scratch[16 + warp]; }
            if (tid < nseg) { carry_incl[tid] = (uint16_t)tr; nseq_s[tid] = (uint8_t)my_nseq; }
        }
        __syncthreads();
        const uint32_t carry_in = (tid > 0 && tid < nseg) ? carry_incl[tid - 1] : 0;

        // ---- P4b: encoded bytes per segment, block scan -> output offsets
        uint32_t my_bytes = 0;
        if (tid < nseg) {
            uint32_t prev_end = seg_start;
            for (uint32_t k = 0; k < my_nseq; k++) {
                const uint2 d = desc[k];
                const uint32_t p = d.x & 0xffff, ml = d.x >> 16;
                const uint32_t ll = (k == 0 ? carry_in : 0) + (p - prev_end);
                my_bytes += 1 + ext_bytes(ll) + ll + 2 + ext_bytes(ml - 4);
                prev_end = p + ml;
            }
        }
        uint32_t total_seq_bytes;
        const uint32_t my_off = block_excl_scan(my_bytes, &total_seq_bytes, scratch);
        if (tid < nseg && my_nseq) {
            const uint32_t p0 = desc[0].x & 0xffff;
            const uint32_t ll0 = carry_in + (p0 - seg_start);
            delta0[tid] = (int32_t)(my_off + 1 + ext_bytes(ll0)) - (int32_t)(p0 - ll0);
        }
        // first later segment that has a sequence (it owns this segment's trailing literals): warp ballots + one barrier
        uint32_t my_next_has = 0xffff;
        {
            const uint32_t hm = __ballot_sync(0xffffffffu, tid < nseg && my_nseq > 0);
            if (lane == 0) scratch[40 + warp] = hm;
            __syncthreads();
            const uint32_t above = lane == 31 ? 0u : (hm >> (lane + 1));
            if (above) my_next_has = tid + (uint32_t)__ffs((int)above);
            else for (uint32_t w = warp + 1; w < LZ_THREADS / 32; w++) { const uint32_t m = scratch[40 + w]; if (m) { my_next_has = w * 32 + (uint32_t)__ffs((int)m) - 1; break; } }
        }
        const uint32_t ll_final = nseg ? carry_incl[nseg - 1] : 0;
        const int32_t delta_final = (int32_t)(total_seq_bytes + 1 + ext_bytes(ll_final)) - (int32_t)(len - ll_final);

        LZ_PHASE(3);
        // ---- P5: emit
        if (tid < nseg) {
            uint8_t* o = out + my_off; uint32_t prev_end = seg_start;
            for (uint32_t k = 0; k < my_nseq; k++) {
                const uint2 d = desc[k];
                const uint32_t p = d.x & 0xffff, ml = d.x >> 16, off = d.y;
                const uint32_t cin = (k == 0 ? carry_in : 0);
                const uint32_t ll = cin + (p - prev_end);
                const uint32_t mt = ml - 4;
                *o++ = (uint8_t)(((ll < 15 ? ll : 15) << 4) | (mt < 15 ? mt : 15));
                if (ll >= 15) o = put_ext(o, ll);
                copy_s2g(o + cin, data_w, prev_end, p - prev_end);
                o += ll;
                *o++ = (uint8_t)off; *o++ = (uint8_t)(off >> 8);
                if (mt >= 15) o = put_ext(o, mt);
                prev_end = p + ml;
            }
            if (seg_end > prev_end) {   // trailing literals belong to the next sequence downstream
                const uint32_t nh = my_next_has;
                const int32_t dl = (nh != 0xffff && nh < nseg) ? delta0[nh] : delta_final;
                copy_s2g(out + ((int32_t)prev_end + dl), data_w, prev_end, seg_end - prev_end);
            }
        }
        if (tid == 0) {
            uint8_t* o = out + total_seq_bytes;
            *o++ = (uint8_t)((ll_final < 15 ? ll_final : 15) << 4);
            if (ll_final >= 15) o = put_ext(o, ll_final);
            a.comp_size[f] = total_seq_bytes + 1 + ext_bytes(ll_final) + ll_final;
        }
        __syncthreads();
        LZ_PHASE(4);
    }
}
#endif  // TF_KERNELS_LZ4

// ------------------------------------------------------------------ CityHash128 v1.0.2 over [method byte .. end of block]
namespace cityd {
#define CK0 0xc3a5c85c97cb3127ULL
#define CK1 0xb492b66fbe98f273ULL
#define CK2 0x9ae16a3b2f90404fULL
#define CK3 0xc949d7c7509e6557ULL
__device__ __forceinline__ uint64_t f64(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p; const uint64_t* q = (const uint64_t*)(a & ~(uintptr_t)7); const uint32_t s = (uint32_t)(a & 7) * 8;
    if (s == 0) return q[0];
    return (q[0] >> s) | (q[1] << (64 - s));
}
__device__ __forceinline__ uint32_t f32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
__device__ __forceinline__ uint64_t rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
__device__ __forceinline__ uint64_t smix(uint64_t v) { return v ^ (v >> 47); }
__device__ __forceinline__ uint64_t hl16(uint64_t u, uint64_t v) {
    const uint64_t kMul = 0x9ddfea08eb382d69ULL;
    uint64_t a = (u ^ v) * kMul; a ^= (a >> 47);
    uint64_t b = (v ^ a) * kMul; b ^= (b >> 47); b *= kMul; return b;
}
__device__ inline uint64_t hl0to16(const uint8_t* s, size_t len) {
    if (len > 8) { uint64_t a = f64(s), b = f64(s + len - 8); return hl16(a, rot(b + len, (int)len)) ^ b; }
    if (len >= 4) { uint64_t a = f32(s); return hl16(len + (a << 3), f32(s + len - 4)); }
    if (len > 0) { uint8_t a = s[0], b = s[len >> 1], c = s[len - 1]; uint32_t y = (uint32_t)a + ((uint32_t)b << 8); uint32_t z = (uint32_t)len + ((uint32_t)c << 2); return smix(y * CK2 ^ z * CK3) * CK2; }
    return CK2;
}
struct P { uint64_t first, second; };
__device__ __forceinline__ P weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a, uint64_t b) {
    a += w; b = rot(b + a + z, 21); const uint64_t c = a; a += x; a += y; b += rot(a, 44); P r; r.first = a + z; r.second = b + c; return r;
}
__device__ __forceinline__ P weak32p(const uint8_t* s, uint64_t a, uint64_t b) { return weak32(f64(s), f64(s + 8), f64(s + 16), f64(s + 24), a, b); }
__device__ inline P murmur(const uint8_t* s, size_t len, P seed) {
    uint64_t a = seed.first, b = seed.second, c = 0, d = 0; long l = (long)len - 16;
    if (l <= 0) { a = smix(a * CK1) * CK1; c = b * CK1 + hl0to16(s, len); d = smix(a + (len >= 8 ? f64(s) : c)); }
    else {
        c = hl16(f64(s + len - 8) + CK1, a); d = hl16(b + len, c + f64(s + len - 16)); a += d;
        do { a ^= smix(f64(s) * CK1) * CK1; a *= CK1; b ^= a; c ^= smix(f64(s + 8) * CK1) * CK1; c *= CK1; d ^= c; s += 16; l -= 16; } while (l > 0);
    }
    a = hl16(a, c); b = hl16(d, b);
    P r; r.first = a ^ b; r.second = hl16(b, a); return r;
}
__device__ inline P hash128_seed(const uint8_t* s, size_t len, P seed) {
    if (len < 128) return murmur(s, len, seed);
    P v, w; uint64_t x = seed.first, y = seed.second, z = len * CK1;
    v.first = rot(y ^ CK1, 49) * CK1 + f64(s);
    v.second = rot(v.first, 42) * CK1 + f64(s + 8);
    w.first = rot(y + z, 35) * CK1 + x;
    w.second = rot(x + f64(s + 88), 53) * CK1;
    do {
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            x = rot(x + y + v.first + f64(s + 16), 37) * CK1;
            y = rot(y + v.second + f64(s + 48), 42) * CK1;
            x ^= w.second; y ^= v.first; z = rot(z ^ w.first, 33);
            v = weak32p(s, v.second * CK1, x + w.first);
            w = weak32p(s + 32, z + w.second, y);
            const uint64_t t = z; z = x; x = t; s += 64;
        }
        len -= 128;
    } while (len >= 128);
    y += rot(w.first, 37) * CK0 + z;
    x += rot(v.first + z, 49) * CK0;
    for (size_t tail = 0; tail < len;) {
        tail += 32;
        y = rot(y - x, 42) * CK0 + v.second;
        w.first += f64(s + len - tail + 16);
        x = rot(x, 49) * CK0 + w.first;
        w.first += v.first;
        v = weak32p(s + len - tail, v.first, v.second);
    }
    x = hl16(x, v.first); y = hl16(y, w.first);
    P r; r.first = hl16(x + v.second, w.second) + y; r.second = hl16(x + w.second, y + v.second); return r;
}
__device__ inline P hash128(const uint8_t* s, size_t len) {
    P seed;
    if (len >= 16) { seed.first = f64(s) ^ CK3; seed.second = f64(s + 8); return hash128_seed(s + 16, len - 16, seed); }
    if (len >= 8) { seed.first = f64(s) ^ (len * CK0); seed.second = f64(s + len - 8) ^ CK1; return hash128_seed(nullptr, 0, seed); }
    seed.first = CK0; seed.second = CK1; return hash128_seed(s, len, seed);
}
}  // namespace cityd

struct FrameArgs { uint8_t* slots; uint32_t slot_stride; const uint32_t* comp_size; DState* st; uint32_t frame_bytes; uint64_t* wire_off; uint8_t* wire;
                   uint64_t* tail; };   // tail[0] = frame count, written by k_frame_scan: the checksum / gather kernels of this batch may still run when the next batch resets DState

// CityHash128 is a serial chain per frame, so the parallelism is ACROSS frames: one thread per frame. What a
// thread-per-frame loop would ruin is the memory access (every lane striding through its own frame), so each
// warp stages the next 512 bytes of all its 32 frames with coalesced 16-byte cp.async copies into shared
// memory (double buffered) while the lanes hash the previous 512 bytes out of it.
#define SEAL_STEP 512
#define SEAL_STRIDE 528
#define SEAL_STAGES 2      /* deeper pipelines do not help: the per-frame hash is a serial dependency chain (~43 k instructions) */
#define SEAL_SMEM ((size_t)SEAL_STAGES * 32 * SEAL_STRIDE + 32 * 304)
__device__ __forceinline__ uint64_t sm64(const uint8_t* base, uint32_t off) { return *(const uint64_t*)(base + off); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#ifdef TF_KERNELS_LZ4
__global__ void __launch_bounds__(32) k_frame_seal(FrameArgs a) {
    // SEAL_STAGES steps of 512 bytes per frame in flight: with one warp per SM nothing else hides the load latency
    extern __shared__ __align__(16) uint8_t seal_smem[];
    uint8_t (*s_buf)[32][SEAL_STRIDE] = (uint8_t (*)[32][SEAL_STRIDE])seal_smem;                                  // [SEAL_STAGES][32][SEAL_STRIDE]
    uint8_t (*s_tail)[304] = (uint8_t (*)[304])(seal_smem + (size_t)SEAL_STAGES * 32 * SEAL_STRIDE);              // [32][304]
    const uint32_t lane = threadIdx.x;
    const uint64_t nf = a.tail[0];
    const uint64_t f = (uint64_t)blockIdx.x * 32 + lane;
    const bool have = f < nf;
    uint8_t* s = a.slots + (have ? f : 0) * a.slot_stride;
    // the 9 header bytes [0x82][cs][rs] at slot + 16 were written by k_frame_scan; the checksum goes straight into the wire
    // buffer (k_wire_gather copies everything after it, concurrently, on the side stream)
    const uint32_t cs = have ? a.comp_size[f] + 9 : 0;
    uint8_t* wck = a.wire + (have ? a.wire_off[f] : 0);
    auto put_checksum = [&](uint64_t lo, uint64_t hi) {
#pragma unroll
        for (int b = 0; b < 8; b++) { wck[b] = (uint8_t)(lo >> (8 * b)); wck[8 + b] = (uint8_t)(hi >> (8 * b)); }
    };
    const uint8_t* H = s + 16;
    const bool big = have && cs >= 16 + 128 + 16;
    if (have && !big) { const cityd::P h = cityd::hash128(H, cs); put_checksum(h.first, h.second); }
    const uint8_t* body = H + 16; const uint32_t len = big ? cs - 16 : 0;
    const uint32_t nblk = len / 128, used = nblk * 128;
    const uint32_t nsteps = (used + SEAL_STEP - 1) / SEAL_STEP;
    uint32_t max_steps = nsteps;
#pragma unroll
    for (int d = 16; d; d >>= 1) { const uint32_t o = __shfl_xor_sync(0xffffffffu, max_steps, d); max_steps = o > max_steps ? o : max_steps; }
    if (max_steps == 0) return;
    const uint32_t tail_base = (len > 272 ? (len - 272) : 0) & ~15u;
    // cooperative staging helpers: lane l copies bytes [16 l, 16 l + 16) of every frame's current 512-byte step
    auto stage_step = [&](uint32_t st, uint32_t bi) {
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const uint8_t* bj = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)body, j);
            const uint32_t uj = __shfl_sync(0xffffffffu, used, j);
            const uint32_t off = st * SEAL_STEP + lane * 16;
            if (off < uj) cp_async16(&s_buf[bi][j][lane * 16], bj + off);
        }
        cp_async_commit();
    };
    for (int j = 0; j < 32; j++) {      // tail windows (<= 288 bytes each)
        const uint8_t* bj = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)body, j);
        const uint32_t lj = __shfl_sync(0xffffffffu, len, j), tj = __shfl_sync(0xffffffffu, tail_base, j);
        if (lane < 18 && lj && tj + 16 * lane < lj + 16) cp_async16(&s_tail[j][16 * lane], bj + tj + 16 * lane);
    }
#pragma unroll
    for (int k = 0; k < SEAL_STAGES - 1; k++) { if ((uint32_t)k < max_steps) stage_step(k, k); else cp_async_commit(); }      // the tail windows ride in the first group
    cityd::P v, w; uint64_t x = 0, y = 0, z = 0;
    if (big) {
        x = cityd::f64(H) ^ CK3; y = cityd::f64(H + 8); z = (uint64_t)len * CK1;
        v.first = cityd::rot(y ^ CK1, 49) * CK1 + cityd::f64(body);
        v.second = cityd::rot(v.first, 42) * CK1 + cityd::f64(body + 8);
        w.first = cityd::rot(y + z, 35) * CK1 + x;
        w.second = cityd::rot(x + cityd::f64(body + 88), 53) * CK1;
    }
    for (uint32_t st = 0; st < max_steps; st++) {
        if (st + SEAL_STAGES - 1 < max_steps) stage_step(st + SEAL_STAGES - 1, (st + SEAL_STAGES - 1) % SEAL_STAGES); else cp_async_commit();
        cp_async_wait<SEAL_STAGES - 1>();      // every group but the newest SEAL_STAGES - 1 has landed: step st is in shared memory
        __syncwarp();
        if (st < nsteps) {
            const uint8_t* b = s_buf[st % SEAL_STAGES][lane];
            const uint32_t nb = (nblk - st * 4 < 4) ? nblk - st * 4 : 4;
            for (uint32_t i = 0; i < 2 * nb; i++) {
                const uint32_t o = i * 64;
                x = cityd::rot(x + y + v.first + sm64(b, o + 16), 37) * CK1;
                y = cityd::rot(y + v.second + sm64(b, o + 48), 42) * CK1;
                x ^= w.second; y ^= v.first; z = cityd::rot(z ^ w.first, 33);
                v = cityd::weak32(sm64(b, o), sm64(b, o + 8), sm64(b, o + 16), sm64(b, o + 24), v.second * CK1, x + w.first);
                w = cityd::weak32(sm64(b, o + 32), sm64(b, o + 40), sm64(b, o + 48), sm64(b, o + 56), z + w.second, y);
                const uint64_t t = z; z = x; x = t;
            }
        }
        __syncwarp();
    }
    if (big) {
        const uint32_t rem = len - used;                 // 0..127 bytes; the tail reads reach back into hashed data
        const uint8_t* tb = s_tail[lane];
        auto t64 = [&](uint32_t off) -> uint64_t {      // unaligned 8 bytes at body + off (off >= tail_base)
            const uint32_t o = off - tail_base; const uint32_t al = o & ~7u, sh = (o & 7) * 8;
            const uint64_t lo = *(const uint64_t*)(tb + al);
            if (!sh) return lo;
            return (lo >> sh) | (*(const uint64_t*)(tb + al + 8) << (64 - sh));
        };
        y += cityd::rot(w.first, 37) * CK0 + z;
        x += cityd::rot(v.first + z, 49) * CK0;
        for (uint32_t td = 0; td < rem;) {
            td += 32;
            const uint32_t base = len - td;
            y = cityd::rot(y - x, 42) * CK0 + v.second;
            w.first += t64(base + 16);
            x = cityd::rot(x, 49) * CK0 + w.first;
            w.first += v.first;
            v = cityd::weak32(t64(base), t64(base + 8), t64(base + 16), t64(base + 24), v.first, v.second);
        }
        x = cityd::hl16(x, v.first); y = cityd::hl16(y, w.first);
        put_checksum(cityd::hl16(x + v.second, w.second) + y, cityd::hl16(x + w.second, y + v.second));
    }
}
#endif  // TF_KERNELS_LZ4

// exclusive scan of frame sizes (single block) -> position of every frame in the wire buffer; also stamps the 9-byte
// [method][sizes] header of every frame, which both the checksum (k_frame_seal) and the gather read
#ifdef TF_KERNELS_LZ4
__global__ void __launch_bounds__(1024) k_frame_scan(FrameArgs a) {
    __shared__ uint32_t sm[33];
    const uint64_t nf = a.st->n_frames;
    uint64_t carry = 0;
    for (uint64_t base = 0; base < nf; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = i < nf ? a.comp_size[i] + LZ_HDR : 0;
        uint32_t tot; const uint32_t ex = block_excl_scan(v, &tot, sm);
        if (i < nf) {
            a.wire_off[i] = carry + ex;
            // bytes 16..31 of the slot as two aligned words: [0x82][compressed size + 9][raw size] + the first 7 bytes of the LZ4 block (kept)
            const uint32_t cs = v - LZ_HDR + 9;
            const uint64_t pos0 = i * a.frame_bytes;
            const uint32_t rs = (uint32_t)((a.st->raw_total - pos0 < a.frame_bytes) ? a.st->raw_total - pos0 : a.frame_bytes);
            uint64_t* q = (uint64_t*)(a.slots + i * a.slot_stride + 16);
            q[0] = 0x82ull | ((uint64_t)cs << 8) | ((uint64_t)(rs & 0xffffff) << 40);
            q[1] = (q[1] & ~0xffull) | (uint64_t)(rs >> 24);
        }
        carry += tot;
    }
    if (threadIdx.x == 0) { a.st->wire_total = carry; a.tail[0] = nf; }
}
#endif  // TF_KERNELS_LZ4

// gather the sealed frames into one contiguous stream; source slots are 16-byte aligned, the destination
// is re-aligned with the same shuffle + funnel-shift trick as k_encode_fixed so stores are aligned words
#ifdef TF_KERNELS_LZ4
__global__ void __launch_bounds__(256) k_wire_gather(FrameArgs a) {
    const uint64_t nf = a.tail[0];
    for (uint64_t f = blockIdx.x; f < nf; f += gridDim.x) {
        const uint32_t* src = (const uint32_t*)(a.slots + f * a.slot_stride);
        const uint32_t total = a.comp_size[f] + LZ_HDR;
        const uint64_t base = a.wire_off[f];
        const uint32_t m = (uint32_t)(base & 3);
        const uint32_t T = (m + total + 3) >> 2;
        uint8_t* dst0 = a.wire + (base - m);
        for (uint32_t t = 4 + threadIdx.x; t < T; t += blockDim.x) {       // words 0..3 hold only checksum bytes (m <= 3: word 3 ends at stream byte <= 15)
            const uint32_t wcur = src[t];                       // slot has >= 8 bytes of slack past the block
            const uint32_t wprev = t ? src[t - 1] : 0;
            const uint32_t val = m ? __funnelshift_r(wprev, wcur, 8 * (4 - m)) : wcur;
            const int32_t sb = (int32_t)(4 * t) - (int32_t)m;
            uint8_t* dst = dst0 + 4 * (uint64_t)t;
            if (sb >= 16 && (uint32_t)sb + 4 <= total) *(uint32_t*)dst = val;          // bytes [0, 16) are the checksum: k_frame_seal writes them
            else {
#pragma unroll
                for (int b = 0; b < 4; b++) { const int32_t x = sb + b; if (x >= 16 && (uint32_t)x < total) dst[b] = (uint8_t)(val >> (8 * b)); }
            }
        }
    }
}
#endif  // TF_KERNELS_LZ4

}  // namespace tfk
