// String columns of the ClickHouse block / the columnar output: sizes, offsets, payload.
#pragma once
#include "device_types.cuh"
#include "kernels_encode.cuh"
#include "kernels_fmt.cuh"

namespace tfk {

// ------------------------------------------------------------------ String columns
__device__ __forceinline__ uint32_t str_len(const DCol& c, const uint32_t* sel, uint64_t j, uint64_t n, uint64_t& r) {
    if (j >= n) { r = 0; return 0xffffffffu; }
    r = sel ? sel[j] : j;
    if (c.out_kind == OK_TOSTR) { CountSink cs; cs.n = 0; fmt_value(cs, c, r); return cs.n; }    // convert_to_string: length of the text form
    if (!row_valid(c, r)) return 0;
    return c.offsets[r + 1] - c.offsets[r];
}

// encoded size (LEB128 length + payload) of every tile of TF_STR_TILE kept rows, for every String column
__global__ void __launch_bounds__(TF_STR_THREADS) k_str_sizes(EncodeArgs a) {
    __shared__ uint32_t sm[33];
    const DCol c = a.cols[a.slots[blockIdx.y]];
    const uint64_t n = a.st->n_kept;
    const uint64_t j0 = (uint64_t)blockIdx.x * TF_STR_TILE;
    if (j0 >= n) return;
    uint64_t r; const uint32_t L = str_len(c, a.sel, j0 + threadIdx.x, n, r);
    uint32_t tot; block_excl_scan(L != 0xffffffffu ? L + (a.columnar ? 0 : varint_len(L)) : 0u, &tot, sm);
    if (threadIdx.x == 0) a.tile_sum[(size_t)c.str_slot * a.ntiles_cap + blockIdx.x] = tot;
}

// LEB128 length + bytes, one kept row per thread; neighbouring threads own neighbouring rows, so a warp reads one
// contiguous span of the source heap (L1 serves the loads). The tile's output (a contiguous span of the block at an
// arbitrary byte address) is assembled in shared memory and written with aligned, coalesced 4-byte stores (same
// funnel-shift re-alignment as k_encode_fixed); tiles larger than the staging buffer (long strings) go direct.
__global__ void __launch_bounds__(TF_STR_THREADS) k_encode_str(EncodeArgs a) {
    __shared__ uint32_t sm[33];
    __shared__ __align__(16) uint8_t stage[TF_STR_STAGE + 8];
    const DCol c = a.cols[a.slots[blockIdx.y]];
    const uint64_t n = a.st->n_kept;
    const uint64_t j0 = (uint64_t)blockIdx.x * TF_STR_TILE;
    if (j0 >= n) return;
    uint64_t R; const uint32_t L = str_len(c, a.sel, j0 + threadIdx.x, n, R);
    const bool tostr = c.out_kind == OK_TOSTR;
    const uint8_t* s = (L != 0xffffffffu && L && !tostr) ? c.heap + c.offsets[R] : nullptr;
    if (tostr && c.nullable && !a.columnar && L != 0xffffffffu) a.raw[c.null_off + j0 + threadIdx.x] = 0;   // "<nil>" is a value
    uint32_t tot; const uint32_t ex = block_excl_scan(L != 0xffffffffu ? L + (a.columnar ? 0 : varint_len(L)) : 0u, &tot, sm);
    const uint64_t tb = a.tile_base[(size_t)c.str_slot * a.ntiles_cap + blockIdx.x];
    uint8_t* gdst = a.raw + c.out_off + tb;
    if (a.columnar && L != 0xffffffffu) ((uint32_t*)(a.raw + c.offs_off))[j0 + threadIdx.x] = (uint32_t)(tb + ex);
    const bool staged = tot <= TF_STR_STAGE;
    uint8_t* o = staged ? stage + ex : gdst + ex;
    if (L != 0xffffffffu) {
        if (!a.columnar) {
            uint32_t v = L;
            while (v >= 0x80) { *o++ = (uint8_t)(v | 0x80); v >>= 7; }
            *o++ = (uint8_t)v;
        }
        // The copy is latency bound if every byte waits for its own load, so each round issues four independent
        // aligned word loads (16 source bytes, re-aligned with funnel shifts) before any byte is stored.
        if (tostr) { MemSink ms; ms.p = o; fmt_value(ms, c, R); o = ms.p; }
        uint32_t nb = tostr ? 0 : L;
        const uint32_t sh = ((uint32_t)(uintptr_t)s & 3) * 8;
        const uint32_t* sw = (const uint32_t*)((uintptr_t)s & ~(uintptr_t)3);
        while (nb) {
            const uint32_t take = nb < 16 ? nb : 16;
            const uint32_t need = (take + (sh >> 3) + 3) >> 2;            // aligned words that hold these bytes (1..5)
            uint32_t w0 = __ldg(sw), w1 = need > 1 ? __ldg(sw + 1) : 0, w2 = need > 2 ? __ldg(sw + 2) : 0, w3 = need > 3 ? __ldg(sw + 3) : 0, w4 = need > 4 ? __ldg(sw + 4) : 0;
            if (sh) { w0 = __funnelshift_r(w0, w1, sh); w1 = __funnelshift_r(w1, w2, sh); w2 = __funnelshift_r(w2, w3, sh); w3 = __funnelshift_r(w3, w4, sh); }
            const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
                for (int b = 0; b < 4; b++) if ((uint32_t)(4 * q + b) < take) o[4 * q + b] = (uint8_t)(ww[q] >> (8 * b));
            }
            o += take; sw += 4; nb -= take;
        }
    }
    if (!staged) return;
    __syncthreads();
    const uint32_t m = (uint32_t)((uintptr_t)gdst & 3);
    const uint32_t T = (m + tot + 3) >> 2;
    uint8_t* dst0 = gdst - m;
    const uint32_t* sw = (const uint32_t*)stage;
    for (uint32_t t = threadIdx.x; t < T; t += TF_STR_THREADS) {
        const uint32_t wcur = sw[t], wprev = t ? sw[t - 1] : 0;
        const uint32_t val = m ? __funnelshift_r(wprev, wcur, 8 * (4 - m)) : wcur;
        const int32_t sb = (int32_t)(4 * t) - (int32_t)m;
        uint8_t* dst = dst0 + 4 * (size_t)t;
        if (sb >= 0 && (uint32_t)sb + 4 <= tot) *(uint32_t*)dst = val;
        else {
#pragma unroll
            for (int b = 0; b < 4; b++) { const int32_t x = sb + b; if (x >= 0 && (uint32_t)x < tot) dst[b] = (uint8_t)(val >> (8 * b)); }
        }
    }
}


}  // namespace tfk
