// Reader of the row image of include/tfgpu_sink.h (one tag byte per value, then its payload): shared by the transposer (host_rows.cu) and
// by tfgpu_sink_push (host_sink.cu: table_splitter reads the split columns of every row).
#pragma once
#include <cstdint>
#include <cstring>
#include "../../include/tfgpu_sink.h"

namespace {

// (table forms of the two switches below: they sit on the per-value path of both transposer passes)
static const uint8_t PAYLOAD_W[18] = {0, 1, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8, 0xff, 0xff, 12, 8, 0xff, 0xff};   // 0xff: length-prefixed
inline uint32_t payload_fixed_slow(int tag) {
    switch (tag) {
    case TF_V_NIL: return 0; case TF_V_BOOL: case TF_V_INT8: case TF_V_UINT8: return 1; case TF_V_INT16: case TF_V_UINT16: return 2;
    case TF_V_INT32: case TF_V_UINT32: case TF_V_FLOAT32: return 4; case TF_V_INT64: case TF_V_UINT64: case TF_V_FLOAT64: case TF_V_DURATION: return 8;
    case TF_V_TIME: return 12; default: return 0xffffffffu;     // length-prefixed
    }
}
inline uint32_t payload_fixed(int tag) { const uint8_t w = PAYLOAD_W[tag]; return w == 0xff ? 0xffffffffu : w; }

// one value of the image: tag, payload pointer, payload length (text length for the length-prefixed tags)
struct Val { int tag; const uint8_t* p; uint32_t n; };
inline bool read_val(const uint8_t*& at, const uint8_t* end, Val& v) {
    if (at >= end) return false;
    v.tag = *at++;
    if (v.tag > TF_V_JSON) return false;
    uint32_t w = payload_fixed(v.tag);
    if (w == 0xffffffffu) { if (end - at < 4) return false; std::memcpy(&w, at, 4); at += 4; }
    if ((size_t)(end - at) < w) return false;
    v.p = at; v.n = w; at += w; return true;
}
inline int64_t val_i64(const Val& v) {
    switch (v.tag) {
    case TF_V_BOOL: return v.p[0] ? 1 : 0;
    case TF_V_INT8: return (int8_t)v.p[0]; case TF_V_UINT8: return v.p[0];
    case TF_V_INT16: { int16_t x; std::memcpy(&x, v.p, 2); return x; } case TF_V_UINT16: { uint16_t x; std::memcpy(&x, v.p, 2); return x; }
    case TF_V_INT32: { int32_t x; std::memcpy(&x, v.p, 4); return x; } case TF_V_UINT32: { uint32_t x; std::memcpy(&x, v.p, 4); return x; }
    default: { int64_t x; std::memcpy(&x, v.p, 8); return x; }
    }
}

}  // namespace
