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

} // namespace rsb200
