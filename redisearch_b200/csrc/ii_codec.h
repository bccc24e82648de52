// One record of an IndexBlock, for every term codec of the reference (RS/inverted_index/src/codec/*.rs), shared by the device
// decoders (ii_kernels.cu) and the host decoder (ii_host.cpp).  Numbering = II_Codec (include/ii_b200.h).
//
//   0 Full              qint4[delta, freq, fieldMask, offsetsLen] + offsets                  full.rs:38-64
//   1 FreqsOnly         qint2[delta, freq]                                                   freqs_only.rs:33
//   2 FreqsFields       qint3[delta, freq, fieldMask]                                        freqs_fields.rs:43
//   3 FieldsOnly        qint2[delta, fieldMask]                                              fields_only.rs:43
//   4 DocIdsOnly        varint(delta)                                                        doc_ids_only.rs:33
//   5 RawDocIdsOnly     u32 LE (docId - block.first_doc_id)                                  raw_doc_ids_only.rs:31-37
//   6 FreqsOffsets      qint3[delta, freq, offsetsLen] + offsets                             freqs_offsets.rs:32-64
//   7 OffsetsOnly       qint2[delta, offsetsLen] + offsets            (freq 1)               offsets_only.rs:31-62
//   8 FieldsOffsets     qint3[delta, fieldMask, offsetsLen] + offsets (freq 1)               fields_offsets.rs:36-84
//   9 FullWide          qint3[delta, freq, offsetsLen] + varint(fieldMask u128) + offsets    full.rs:197-232
//  10 FreqsFieldsWide   qint2[delta, freq] + varint(fieldMask u128)                          freqs_fields.rs:114-145
//  11 FieldsOnlyWide    varint(delta) + varint(fieldMask u128)                               fields_only.rs:109-137
//  12 FieldsOffsetsWide qint2[delta, offsetsLen] + varint(fieldMask u128) + offsets (freq 1) fields_offsets.rs:138-185
//
// qint: one lead byte, 2 bits per value = its byte length - 1, values little endian (RS/qint/src/lib.rs:149-286).
// varint: 7-bit groups, most significant first, +1 per continuation (RS/varint/src/lib.rs:113-200).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define II_HD __host__ __device__ __forceinline__
#else
#define II_HD inline
#endif

namespace rsb200 {

constexpr int kNumCodecs = 13;
II_HD bool ii_codec_has_offsets(int c) { return c == 0 || c == 6 || c == 7 || c == 8 || c == 9 || c == 12; }
II_HD bool ii_codec_has_mask(int c) { return c == 0 || c == 2 || c == 3 || c == 8 || c >= 9; }
II_HD bool ii_codec_is_wide(int c) { return c >= 9 && c <= 12; }

struct IIRecord {
    uint64_t delta;           // from the previous docId of the block (codec 5: from the block's first docId)
    uint32_t freq;            // 1 for the codecs that do not store it
    uint64_t mask_lo, mask_hi; // all ones for the codecs without a field mask (RS_FIELDMASK_ALL)
    uint32_t off_len;         // bytes of the offsets payload
    const uint8_t *offsets;   // where it starts (the next record follows it)
};

II_HD uint32_t ii_le(const uint8_t *p, int nb) {
    uint32_t v = p[0];
    if (nb > 1) v |= (uint32_t)p[1] << 8;
    if (nb > 2) v |= (uint32_t)p[2] << 16;
    if (nb > 3) v |= (uint32_t)p[3] << 24;
    return v;
}

// Decode the record at p.  Returns the address of the NEXT record, or nullptr when the record runs past `end` (checked only when
// kChecked; the device decoders trust the block table the host has validated against the byte lengths).
template <bool kChecked>
II_HD const uint8_t *ii_decode_record(const uint8_t *p, const uint8_t *end, int codec, IIRecord &r) {
    r.freq = 1;
    r.mask_lo = r.mask_hi = ~0ull;
    r.off_len = 0;
    r.offsets = nullptr;
#define II_NEED(n)                                    \
    do {                                              \
        if (kChecked && (p + (n) > end)) return nullptr; \
    } while (0)
    auto varint64 = [&](uint64_t &lo, uint64_t &hi) -> bool { // up to 128 bits
        if (kChecked && p >= end) return false;
        uint8_t c = *p++;
        lo = c & 0x7f;
        hi = 0;
        while (c & 0x80) {
            if (kChecked && p >= end) return false;
            // val += 1; val = (val << 7) | next
            lo += 1;
            if (lo == 0) hi += 1;
            c = *p++;
            hi = (hi << 7) | (lo >> 57);
            lo = (lo << 7) | (c & 0x7f);
        }
        return true;
    };
    int nq = 0; // qint values
    switch (codec) {
    case 0: nq = 4; break;
    case 2:
    case 6:
    case 8:
    case 9: nq = 3; break;
    case 1:
    case 3:
    case 7:
    case 10:
    case 12: nq = 2; break;
    default: break;
    }
    uint32_t v[4] = {0, 0, 0, 0};
    if (nq) {
        II_NEED(1);
        const uint8_t lead = *p++;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < nq) {
                const int nb = ((lead >> (2 * i)) & 3) + 1;
                II_NEED(nb);
                v[i] = ii_le(p, nb);
                p += nb;
            }
        r.delta = v[0];
    }
    switch (codec) {
    case 0:
        r.freq = v[1];
        r.mask_lo = v[2];
        r.mask_hi = 0;
        r.off_len = v[3];
        break;
    case 1: r.freq = v[1]; break;
    case 2:
        r.freq = v[1];
        r.mask_lo = v[2];
        r.mask_hi = 0;
        break;
    case 3:
        r.mask_lo = v[1];
        r.mask_hi = 0;
        break;
    case 4: {
        uint64_t lo, hi;
        if (!varint64(lo, hi)) return nullptr;
        r.delta = lo;
        break;
    }
    case 5:
        II_NEED(4);
        r.delta = ii_le(p, 4);
        p += 4;
        break;
    case 6:
        r.freq = v[1];
        r.off_len = v[2];
        break;
    case 7: r.off_len = v[1]; break;
    case 8:
        r.mask_lo = v[1];
        r.mask_hi = 0;
        r.off_len = v[2];
        break;
    case 9:
        r.freq = v[1];
        r.off_len = v[2];
        if (!varint64(r.mask_lo, r.mask_hi)) return nullptr;
        break;
    case 10:
        r.freq = v[1];
        if (!varint64(r.mask_lo, r.mask_hi)) return nullptr;
        break;
    case 11: {
        uint64_t lo, hi;
        if (!varint64(lo, hi)) return nullptr;
        r.delta = lo;
        if (!varint64(r.mask_lo, r.mask_hi)) return nullptr;
        break;
    }
    case 12:
        r.off_len = v[1];
        if (!varint64(r.mask_lo, r.mask_hi)) return nullptr;
        break;
    default: return nullptr;
    }
    r.offsets = p;
    II_NEED(r.off_len);
    return p + r.off_len;
#undef II_NEED
}

// One record of a NUMERIC index block (RS/inverted_index/src/codec/numeric.rs:553-626): header byte = (type-specific << 5) |
// (type << 3) | delta_bytes; delta (0-7 bytes LE); value: TINY 0-7 in the header, INT_POS / INT_NEG a 1-8 byte magnitude, FLOAT an
// f32 / f64 magnitude with the sign in the header, or an infinity.  Returns the next record (nullptr past `end` when kChecked).
template <bool kChecked>
II_HD const uint8_t *ii_decode_numeric(const uint8_t *p, const uint8_t *end, uint64_t &delta, double &value) {
    if (kChecked && p >= end) return nullptr;
    const uint8_t header = *p++;
    const int nd = header & 7;
    const int type = (header >> 3) & 3, upper = header >> 5;
    int nv = 0;
    if (type == 2 || type == 3)
        nv = upper + 1;
    else if (type == 1)
        nv = (upper == 0 || upper == 2) ? 4 : (upper == 4 || upper == 6) ? 8 : 0;
    if (kChecked && p + nd + nv > end) return nullptr;
    uint64_t d = 0;
    for (int i = 0; i < nd; i++) d |= (uint64_t)p[i] << (8 * i);
    p += nd;
    delta = d;
    uint64_t m = 0;
    for (int i = 0; i < nv; i++) m |= (uint64_t)p[i] << (8 * i);
    p += nv;
    if (type == 0) {
        value = (double)upper;
    } else if (type == 2) {
        value = (double)m;
    } else if (type == 3) {
        value = -(double)m; // (num as f64).copysign(-1.0): the magnitude is non-negative
    } else if (nv == 4) {
        union { uint32_t u; float f; } c;
        c.u = (uint32_t)m;
        const double a = (double)c.f;
        value = upper == 2 ? -(a < 0 ? -a : a) : a;
    } else if (nv == 8) {
        union { uint64_t u; double f; } c;
        c.u = m;
        const double a = c.f;
        value = upper == 6 ? -(a < 0 ? -a : a) : a;
    } else {
        union { uint64_t u; double f; } c;
        c.u = (upper == 1 || upper == 5) ? 0x7FF0000000000000ull : 0xFFF0000000000000ull; // +inf / -inf
        value = c.f;
    }
    return p;
}
// NumericFilter::value_in_range, RS/inverted_index/src/reader/numeric.rs:80-85
II_HD bool ii_numeric_in_range(double value, double min, double max, bool min_inclusive, bool max_inclusive) {
    const bool min_ok = value > min || (min_inclusive && value == min);
    const bool max_ok = value < max || (max_inclusive && value == max);
    return min_ok && max_ok;
}

// ---- encoders (host side: the index writer, II_IndexWriter) — the inverse of the decoders above ---------------------------
// qint (RS/qint/src/lib.rs:149-215): lead byte + each value in its minimal number of little-endian bytes (at least one)
inline size_t ii_qint_encode(const uint32_t *vals, int n, uint8_t *out) {
    uint8_t lead = 0;
    size_t pos = 1;
    for (int i = 0; i < n; i++) {
        uint32_t v = vals[i];
        int bytes = 0;
        do {
            out[pos++] = (uint8_t)v;
            bytes++;
            v >>= 8;
        } while (v);
        lead |= (uint8_t)((bytes - 1) << (2 * i));
    }
    out[0] = lead;
    return pos;
}
// varint of up to 128 bits (RS/varint/src/lib.rs:113-160): 7-bit groups, most significant first, minus one per continuation
inline size_t ii_varint_encode(uint64_t lo, uint64_t hi, uint8_t *out) {
    uint8_t buf[24];
    int pos = 23;
    unsigned __int128 v = ((unsigned __int128)hi << 64) | lo;
    buf[pos] = (uint8_t)(v & 0x7f);
    v >>= 7;
    while (v) {
        pos--;
        v -= 1;
        buf[pos] = (uint8_t)(0x80 | (uint8_t)(v & 0x7f));
        v >>= 7;
    }
    for (int i = pos; i < 24; i++) out[i - pos] = buf[i];
    return (size_t)(24 - pos);
}
// one term record; `out` needs 48 + off_len bytes.  The narrow codecs take the low 32 bits of the mask.
inline size_t ii_encode_record(int codec, uint32_t delta, uint32_t freq, uint64_t mask_lo, uint64_t mask_hi, const uint8_t *offsets,
                               uint32_t off_len, uint8_t *out) {
    size_t n = 0;
    const uint32_t m32 = (uint32_t)mask_lo;
    bool with_offsets = false;
    switch (codec) {
    case 0: { const uint32_t v[4] = {delta, freq, m32, off_len}; n = ii_qint_encode(v, 4, out); with_offsets = true; break; }
    case 1: { const uint32_t v[2] = {delta, freq}; n = ii_qint_encode(v, 2, out); break; }
    case 2: { const uint32_t v[3] = {delta, freq, m32}; n = ii_qint_encode(v, 3, out); break; }
    case 3: { const uint32_t v[2] = {delta, m32}; n = ii_qint_encode(v, 2, out); break; }
    case 4: n = ii_varint_encode(delta, 0, out); break;
    case 5: out[0] = (uint8_t)delta, out[1] = (uint8_t)(delta >> 8), out[2] = (uint8_t)(delta >> 16), out[3] = (uint8_t)(delta >> 24); n = 4; break;
    case 6: { const uint32_t v[3] = {delta, freq, off_len}; n = ii_qint_encode(v, 3, out); with_offsets = true; break; }
    case 7: { const uint32_t v[2] = {delta, off_len}; n = ii_qint_encode(v, 2, out); with_offsets = true; break; }
    case 8: { const uint32_t v[3] = {delta, m32, off_len}; n = ii_qint_encode(v, 3, out); with_offsets = true; break; }
    case 9: { const uint32_t v[3] = {delta, freq, off_len}; n = ii_qint_encode(v, 3, out); n += ii_varint_encode(mask_lo, mask_hi, out + n); with_offsets = true; break; }
    case 10: { const uint32_t v[2] = {delta, freq}; n = ii_qint_encode(v, 2, out); n += ii_varint_encode(mask_lo, mask_hi, out + n); break; }
    case 11: n = ii_varint_encode(delta, 0, out); n += ii_varint_encode(mask_lo, mask_hi, out + n); break;
    case 12: { const uint32_t v[2] = {delta, off_len}; n = ii_qint_encode(v, 2, out); n += ii_varint_encode(mask_lo, mask_hi, out + n); with_offsets = true; break; }
    default: return 0;
    }
    if (with_offsets) {
        for (uint32_t i = 0; i < off_len; i++) out[n + i] = offsets[i];
        n += off_len;
    }
    return n;
}
// one numeric record (RS/inverted_index/src/codec/numeric.rs: Value::from :790-838, encode_value :363-520); out >= 17 bytes
inline size_t ii_encode_numeric(uint64_t delta, double value, bool compress_floats, uint8_t *out) {
    auto trim = [](uint64_t v, uint8_t *dst) {
        size_t n = 0;
        while (v) {
            dst[n++] = (uint8_t)v;
            v >>= 8;
        }
        return n;
    };
    uint8_t d[8], v[8];
    const size_t nd = trim(delta, d);
    size_t nv = 0;
    uint8_t type = 0, upper = 0;
    union { double f; uint64_t u; } cv;
    cv.f = value;
    const bool neg = (cv.u >> 63) != 0;
    const double a = neg ? -value : value; // |value| (NaN stays NaN)
    uint64_t u = 0; // `abs as u64` saturates: NaN -> 0, >= 2^64 -> u64::MAX
    if (a == a) u = a >= 18446744073709551616.0 ? ~0ull : (uint64_t)a;
    const bool is_inf = a > 1.7976931348623157e308;
    if ((double)u == a) {
        if ((double)(u & 7) == value) {
            type = 0, upper = (uint8_t)u; // TINY (also -0.0)
        } else {
            type = neg ? 3 : 2;
            nv = trim(u, v);
            upper = (uint8_t)(nv - 1);
        }
    } else if (is_inf) {
        type = 1, upper = neg ? 3 : 1;
    } else {
        type = 1;
        const float f32 = (float)a;
        const double diff = a - (double)f32;
        if ((double)f32 == a || (compress_floats && (diff < 0 ? -diff : diff) < 0.01)) {
            if (f32 == 0.0f) {
                type = 0, upper = 0;
            } else {
                upper = neg ? 2 : 0;
                union { float f; uint32_t w; } c32;
                c32.f = f32;
                v[0] = (uint8_t)c32.w, v[1] = (uint8_t)(c32.w >> 8), v[2] = (uint8_t)(c32.w >> 16), v[3] = (uint8_t)(c32.w >> 24);
                nv = 4;
            }
        } else {
            upper = neg ? 6 : 4;
            union { double f; uint64_t w; } c64;
            c64.f = a;
            for (int i = 0; i < 8; i++) v[i] = (uint8_t)(c64.w >> (8 * i));
            nv = 8;
        }
    }
    out[0] = (uint8_t)((upper << 5) | (type << 3) | (uint8_t)nd);
    for (size_t i = 0; i < nd; i++) out[1 + i] = d[i];
    for (size_t i = 0; i < nv; i++) out[1 + nd + i] = v[i];
    return 1 + nd + nv;
}

} // namespace rsb200
