/* postings_oracle.c — CPU restatement (plain C) of the reference's posting-list path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference implements this half in Rust and there is no Rust toolchain in the build
 * container, so it cannot be executed here; this file follows the Rust sources function by
 * function and is pinned by the reference's own golden byte vectors and known answers,
 * transcribed in tests/golden/postings_golden.json (tests/test_oracle_postings.py):
 *   qint      src/redisearch_rs/qint/src/lib.rs:149-286            (tests: qint/tests/qint.rs)
 *   varint    src/redisearch_rs/varint/src/lib.rs                   (tests: varint/tests/varint.rs)
 *   codecs    src/redisearch_rs/inverted_index/src/codec/{full,freqs_only,freqs_fields,fields_only,
 *             doc_ids_only,raw_doc_ids_only,freqs_offsets,offsets_only,fields_offsets}.rs incl. the
 *             *Wide variants (u128 field masks as varints)            (tests: tests/integration/codec/)
 *   index     src/redisearch_rs/inverted_index/src/index/core.rs:235-358
 *   reader    src/redisearch_rs/inverted_index/src/reader/core.rs:245-345,391-440
 *   leaf      src/redisearch_rs/rqe_iterators/src/inverted_index/core.rs:237-350
 *   AND       src/redisearch_rs/rqe_iterators/src/intersection.rs:103-169,245-339,428-506
 *   OR        src/redisearch_rs/rqe_iterators/src/union_flat.rs:218-524
 *   idf       src/redisearch_rs/idf/src/lib.rs:36-110               (tests: idf/tests/tests.rs)
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ qint / varint ----------- */
size_t orc_qint_encode(const uint32_t *vals, int n, uint8_t *out) {
    uint8_t leading = 0;
    size_t pos = 1;
    for (int i = 0; i < n; i++) {
        uint32_t v = vals[i];
        int bytes = 0;
        do { /* qint_encode_stepwise: at least one byte, little endian, minimal length */
            out[pos++] = (uint8_t)v;
            bytes++;
            v >>= 8;
        } while (v);
        leading |= (uint8_t)((bytes - 1) << (i * 2));
    }
    out[0] = leading;
    return pos;
}
size_t orc_qint_decode(const uint8_t *in, int n, uint32_t *vals) {
    const uint8_t leading = in[0];
    size_t pos = 1;
    for (int i = 0; i < n; i++) {
        const int bytes = ((leading >> (i * 2)) & 3) + 1;
        uint32_t v = 0;
        for (int b = 0; b < bytes; b++) v |= (uint32_t)in[pos + b] << (8 * b);
        vals[i] = v;
        pos += (size_t)bytes;
    }
    return pos;
}
size_t orc_varint_encode(uint64_t v, uint8_t *out) {
    uint8_t buf[16];
    int pos = 15;
    buf[pos] = (uint8_t)(v & 0x7f);
    v >>= 7;
    while (v) {
        pos--;
        v -= 1;
        buf[pos] = (uint8_t)(0x80 | (v & 0x7f));
        v >>= 7;
    }
    memcpy(out, buf + pos, (size_t)(16 - pos));
    return (size_t)(16 - pos);
}
size_t orc_varint_decode(const uint8_t *in, uint64_t *v) {
    size_t pos = 0;
    uint8_t c = in[pos++];
    uint64_t val = c & 0x7f;
    while (c & 0x80) {
        val += 1;
        c = in[pos++];
        val = (val << 7) | (c & 0x7f);
    }
    *v = val;
    return pos;
}

/* varint/src/lib.rs impl_encode!(u128, 19): the same scheme over 128 bits (field masks of the *Wide codecs) */
typedef unsigned __int128 orc_u128;
static size_t varint_encode_u128(orc_u128 v, uint8_t *out) {
    uint8_t buf[24];
    int pos = 23;
    buf[pos] = (uint8_t)(v & 0x7f);
    v >>= 7;
    while (v) {
        pos--;
        v -= 1;
        buf[pos] = (uint8_t)(0x80 | (uint8_t)(v & 0x7f));
        v >>= 7;
    }
    memcpy(out, buf + pos, (size_t)(24 - pos));
    return (size_t)(24 - pos);
}
static size_t varint_decode_u128(const uint8_t *in, orc_u128 *v) {
    size_t pos = 0;
    uint8_t c = in[pos++];
    orc_u128 val = c & 0x7f;
    while (c & 0x80) {
        val += 1;
        c = in[pos++];
        val = (val << 7) | (c & 0x7f);
    }
    *v = val;
    return pos;
}

/* ------------------------------------------------------------------ inverted index ---------- */
typedef struct {
    uint64_t first_id, last_id;
    uint16_t n;
    uint8_t *buf;
    size_t len, cap;
} Block;

struct OrcInvIndex {
    int codec;
    Block *blocks;
    size_t nblocks, cap;
    uint32_t unique_docs;
};

static uint16_t block_entries(int codec) { /* codec/mod.rs:69, doc_ids_only.rs:26, raw_doc_ids_only.rs:24 */
    return (codec == ORC_CODEC_DOCIDS_ONLY || codec == ORC_CODEC_RAW_DOCIDS_ONLY) ? 1000 : 100;
}

OrcInvIndex *orc_ii_new(int codec) {
    OrcInvIndex *ii = (OrcInvIndex *)calloc(1, sizeof(*ii));
    ii->codec = codec;
    return ii;
}
void orc_ii_free(OrcInvIndex *ii) {
    if (!ii) return;
    for (size_t i = 0; i < ii->nblocks; i++) free(ii->blocks[i].buf);
    free(ii->blocks);
    free(ii);
}
size_t orc_ii_num_blocks(const OrcInvIndex *ii) { return ii->nblocks; }
size_t orc_ii_num_docs(const OrcInvIndex *ii) { return ii->unique_docs; }
void orc_ii_block(const OrcInvIndex *ii, size_t b, uint64_t *first_id, uint64_t *last_id, uint16_t *num_entries,
                  const uint8_t **buf, size_t *len) {
    const Block *bl = &ii->blocks[b];
    *first_id = bl->first_id;
    *last_id = bl->last_id;
    *num_entries = bl->n;
    *buf = bl->buf;
    *len = bl->len;
}

static void block_reserve(Block *b, size_t extra) {
    if (b->len + extra > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 64;
        while (nc < b->len + extra) nc *= 2;
        b->buf = (uint8_t *)realloc(b->buf, nc);
        b->cap = nc;
    }
}
static Block *push_block(OrcInvIndex *ii, uint64_t doc_id) {
    if (ii->nblocks == ii->cap) {
        ii->cap = ii->cap ? ii->cap * 2 : 4;
        ii->blocks = (Block *)realloc(ii->blocks, ii->cap * sizeof(Block));
    }
    Block *b = &ii->blocks[ii->nblocks++];
    memset(b, 0, sizeof(*b));
    b->first_id = b->last_id = doc_id; /* IndexBlock::new */
    return b;
}

size_t orc_ii_add(OrcInvIndex *ii, uint64_t doc_id, uint32_t freq, uint32_t field_mask, const uint8_t *offsets,
                  uint32_t offsets_len) {
    return orc_ii_add_wide(ii, doc_id, freq, field_mask, 0, offsets, offsets_len);
}
size_t orc_ii_add_wide(OrcInvIndex *ii, uint64_t doc_id, uint32_t freq, uint64_t mask_lo, uint64_t mask_hi, const uint8_t *offsets,
                       uint32_t offsets_len) {
    const uint32_t field_mask = (uint32_t)mask_lo; /* the narrow codecs take a u32 mask (the reference panics on a wider one) */
    const orc_u128 wide_mask = ((orc_u128)mask_hi << 64) | mask_lo;
    /* index/core.rs:244-256: none of the codecs here allow duplicates -> a repeated docId is dropped */
    if (ii->nblocks && ii->blocks[ii->nblocks - 1].last_id == doc_id) return 0;
    /* take_block (:339-358) */
    Block *b;
    if (ii->nblocks == 0 || ii->blocks[ii->nblocks - 1].n >= block_entries(ii->codec))
        b = push_block(ii, doc_id);
    else
        b = &ii->blocks[ii->nblocks - 1];
    /* delta base: previous docId, or the block's first docId for RawDocIdsOnly (raw_doc_ids_only.rs:40-47) */
    uint64_t base = (ii->codec == ORC_CODEC_RAW_DOCIDS_ONLY) ? b->first_id : b->last_id;
    uint64_t delta64 = doc_id - base;
    if (delta64 > 0xFFFFFFFFull) { /* :272-285: delta does not fit -> fresh block, delta 0 */
        b = push_block(ii, doc_id);
        delta64 = 0;
    }
    const uint32_t delta = (uint32_t)delta64;
    const size_t before = b->len;
    block_reserve(b, 64 + offsets_len);
    uint8_t *w = b->buf + b->len;
    size_t nw = 0;
    switch (ii->codec) {
    case ORC_CODEC_FULL: {
        uint32_t v[4] = {delta, freq, field_mask, offsets_len};
        nw = orc_qint_encode(v, 4, w);
        if (offsets_len) memcpy(w + nw, offsets, offsets_len);
        nw += offsets_len;
        break;
    }
    case ORC_CODEC_FREQS_ONLY: {
        uint32_t v[2] = {delta, freq};
        nw = orc_qint_encode(v, 2, w);
        break;
    }
    case ORC_CODEC_FREQS_FIELDS: {
        uint32_t v[3] = {delta, freq, field_mask};
        nw = orc_qint_encode(v, 3, w);
        break;
    }
    case ORC_CODEC_FIELDS_ONLY: {
        uint32_t v[2] = {delta, field_mask};
        nw = orc_qint_encode(v, 2, w);
        break;
    }
    case ORC_CODEC_DOCIDS_ONLY: nw = orc_varint_encode(delta, w); break;
    case ORC_CODEC_RAW_DOCIDS_ONLY: memcpy(w, &delta, 4); nw = 4; break;
    case ORC_CODEC_FREQS_OFFSETS: { /* freqs_offsets.rs:32-50 */
        uint32_t v[3] = {delta, freq, offsets_len};
        nw = orc_qint_encode(v, 3, w);
        if (offsets_len) memcpy(w + nw, offsets, offsets_len);
        nw += offsets_len;
        break;
    }
    case ORC_CODEC_OFFSETS_ONLY: { /* offsets_only.rs:31-48 */
        uint32_t v[2] = {delta, offsets_len};
        nw = orc_qint_encode(v, 2, w);
        if (offsets_len) memcpy(w + nw, offsets, offsets_len);
        nw += offsets_len;
        break;
    }
    case ORC_CODEC_FIELDS_OFFSETS: { /* fields_offsets.rs:36-60 */
        uint32_t v[3] = {delta, field_mask, offsets_len};
        nw = orc_qint_encode(v, 3, w);
        if (offsets_len) memcpy(w + nw, offsets, offsets_len);
        nw += offsets_len;
        break;
    }
    case ORC_CODEC_FULL_WIDE: { /* full.rs:197-217 */
        uint32_t v[3] = {delta, freq, offsets_len};
        nw = orc_qint_encode(v, 3, w);
        nw += varint_encode_u128(wide_mask, w + nw);
        if (offsets_len) memcpy(w + nw, offsets, offsets_len);
        nw += offsets_len;
        break;
    }
    case ORC_CODEC_FREQS_FIELDS_WIDE: { /* freqs_fields.rs:114-126 */
        uint32_t v[2] = {delta, freq};
        nw = orc_qint_encode(v, 2, w);
        nw += varint_encode_u128(wide_mask, w + nw);
        break;
    }
    case ORC_CODEC_FIELDS_ONLY_WIDE: /* fields_only.rs:109-121: two varints */
        nw = orc_varint_encode(delta, w);
        nw += varint_encode_u128(wide_mask, w + nw);
        break;
    case ORC_CODEC_FIELDS_OFFSETS_WIDE: { /* fields_offsets.rs:138-160 */
        uint32_t v[2] = {delta, offsets_len};
        nw = orc_qint_encode(v, 2, w);
        nw += varint_encode_u128(wide_mask, w + nw);
        if (offsets_len) memcpy(w + nw, offsets, offsets_len);
        nw += offsets_len;
        break;
    }
    }
    b->len += nw;
    b->n++;
    b->last_id = doc_id;
    ii->unique_docs++;
    return b->len - before;
}

/* ------------------------------------------------------------------ reader ------------------ */
struct OrcReader {
    const OrcInvIndex *ii;
    orc_u128 mask;
    size_t cur_block, buf_pos;
    uint64_t last_doc_id;
    uint16_t entry_in_block;
};

static void set_block(OrcReader *r, size_t idx) { /* reader/core.rs:430-440 */
    r->cur_block = idx;
    r->last_doc_id = r->ii->blocks[idx].first_id;
    r->buf_pos = 0;
    r->entry_in_block = 0;
}
OrcReader *orc_reader_new(const OrcInvIndex *ii, uint32_t field_mask_filter) {
    return orc_reader_new_wide(ii, field_mask_filter, 0);
}
OrcReader *orc_reader_new_wide(const OrcInvIndex *ii, uint64_t filter_lo, uint64_t filter_hi) {
    OrcReader *r = (OrcReader *)calloc(1, sizeof(*r));
    r->ii = ii;
    r->mask = ((orc_u128)filter_hi << 64) | filter_lo;
    orc_reader_rewind(r);
    return r;
}
void orc_reader_free(OrcReader *r) { free(r); }
void orc_reader_rewind(OrcReader *r) {
    r->cur_block = 0;
    r->buf_pos = 0;
    r->entry_in_block = 0;
    r->last_doc_id = r->ii->nblocks ? r->ii->blocks[0].first_id : 0;
}

/* decode one record at the cursor; base = previous docId (or block first id for raw ids) */
static void decode_at(const OrcReader *r, const Block *b, size_t *pos, uint64_t base, uint64_t *doc_id, uint32_t *freq,
                      orc_u128 *mask) {
    const uint8_t *in = b->buf + *pos;
    uint32_t v[4];
    uint64_t u;
    *freq = 1;
    *mask = ~(orc_u128)0; /* codecs without a mask match every field (RS_FIELDMASK_ALL) */
    switch (r->ii->codec) {
    case ORC_CODEC_FULL:
        *pos += orc_qint_decode(in, 4, v);
        *doc_id = base + v[0];
        *freq = v[1];
        *mask = v[2];
        *pos += v[3];
        break;
    case ORC_CODEC_FREQS_ONLY:
        *pos += orc_qint_decode(in, 2, v);
        *doc_id = base + v[0];
        *freq = v[1];
        break;
    case ORC_CODEC_FREQS_FIELDS:
        *pos += orc_qint_decode(in, 3, v);
        *doc_id = base + v[0];
        *freq = v[1];
        *mask = v[2];
        break;
    case ORC_CODEC_FIELDS_ONLY:
        *pos += orc_qint_decode(in, 2, v);
        *doc_id = base + v[0];
        *mask = v[1];
        break;
    case ORC_CODEC_DOCIDS_ONLY:
        *pos += orc_varint_decode(in, &u);
        *doc_id = base + u;
        break;
    case ORC_CODEC_RAW_DOCIDS_ONLY: {
        uint32_t d;
        memcpy(&d, in, 4);
        *pos += 4;
        *doc_id = b->first_id + d;
        break;
    }
    case ORC_CODEC_FREQS_OFFSETS: /* freqs_offsets.rs:52-64 */
        *pos += orc_qint_decode(in, 3, v);
        *doc_id = base + v[0];
        *freq = v[1];
        *pos += v[2];
        break;
    case ORC_CODEC_OFFSETS_ONLY: /* offsets_only.rs:50-62: freq 1 */
        *pos += orc_qint_decode(in, 2, v);
        *doc_id = base + v[0];
        *pos += v[1];
        break;
    case ORC_CODEC_FIELDS_OFFSETS: /* fields_offsets.rs:62-84: freq 1 */
        *pos += orc_qint_decode(in, 3, v);
        *doc_id = base + v[0];
        *mask = v[1];
        *pos += v[2];
        break;
    case ORC_CODEC_FULL_WIDE: { /* full.rs:219-232 */
        size_t q = orc_qint_decode(in, 3, v);
        q += varint_decode_u128(in + q, mask);
        *pos += q + v[2];
        *doc_id = base + v[0];
        *freq = v[1];
        break;
    }
    case ORC_CODEC_FREQS_FIELDS_WIDE: { /* freqs_fields.rs:128-145 */
        size_t q = orc_qint_decode(in, 2, v);
        q += varint_decode_u128(in + q, mask);
        *pos += q;
        *doc_id = base + v[0];
        *freq = v[1];
        break;
    }
    case ORC_CODEC_FIELDS_ONLY_WIDE: { /* fields_only.rs:123-137 */
        size_t q = orc_varint_decode(in, &u);
        q += varint_decode_u128(in + q, mask);
        *pos += q;
        *doc_id = base + u;
        break;
    }
    case ORC_CODEC_FIELDS_OFFSETS_WIDE: { /* fields_offsets.rs:162-185: freq 1 */
        size_t q = orc_qint_decode(in, 2, v);
        q += varint_decode_u128(in + q, mask);
        *pos += q + v[1];
        *doc_id = base + v[0];
        break;
    }
    }
}

static int next_unfiltered(OrcReader *r, uint64_t *doc_id, uint32_t *freq, orc_u128 *mask) { /* :245-277 */
    const OrcInvIndex *ii = r->ii;
    if (ii->nblocks == 0) return 0;
    if (ii->blocks[r->cur_block].len <= r->buf_pos) {
        if (r->cur_block + 1 >= ii->nblocks) return 0;
        set_block(r, r->cur_block + 1);
    }
    const Block *b = &ii->blocks[r->cur_block];
    decode_at(r, b, &r->buf_pos, r->last_doc_id, doc_id, freq, mask);
    r->entry_in_block++;
    r->last_doc_id = *doc_id;
    return 1;
}

static int reader_next128(OrcReader *r, uint64_t *doc_id, uint32_t *freq, orc_u128 *mask) {
    for (;;) {
        if (!next_unfiltered(r, doc_id, freq, mask)) return 0;
        if (r->mask == 0 || (*mask & r->mask)) return 1; /* reader/field_mask.rs */
    }
}
int orc_reader_next(OrcReader *r, uint64_t *doc_id, uint32_t *freq, uint32_t *field_mask) {
    orc_u128 m;
    const int ok = reader_next128(r, doc_id, freq, &m);
    if (ok) *field_mask = (uint32_t)m;
    return ok;
}
int orc_reader_next_wide(OrcReader *r, uint64_t *doc_id, uint32_t *freq, uint64_t *mask_lo, uint64_t *mask_hi) {
    orc_u128 m;
    const int ok = reader_next128(r, doc_id, freq, &m);
    if (ok) {
        *mask_lo = (uint64_t)m;
        *mask_hi = (uint64_t)(m >> 64);
    }
    return ok;
}

static int skip_to_block(OrcReader *r, uint64_t target) { /* :309-345 */
    const OrcInvIndex *ii = r->ii;
    if (ii->nblocks == 0) return 0;
    if (ii->blocks[r->cur_block].last_id >= target) return 1;
    if (ii->blocks[ii->nblocks - 1].last_id < target) return 0;
    size_t start = r->cur_block + 1;
    if (start < ii->nblocks && ii->blocks[start].last_id >= target) {
        set_block(r, start);
        return 1;
    }
    size_t lo = start, hi = ii->nblocks; /* binary search by last_doc_id, insertion point on miss */
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (ii->blocks[mid].last_id < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    set_block(r, lo);
    return 1;
}

int orc_reader_seek(OrcReader *r, uint64_t target, uint64_t *doc_id, uint32_t *freq, uint32_t *field_mask) {
    if (!skip_to_block(r, target)) return 0; /* :279-306 */
    const Block *b = &r->ii->blocks[r->cur_block];
    orc_u128 m = 0;
    for (;;) { /* Decoder::seek: decode forward until docId >= target */
        if (b->len <= r->buf_pos) return 0;
        decode_at(r, b, &r->buf_pos, r->last_doc_id, doc_id, freq, &m);
        r->entry_in_block++;
        r->last_doc_id = *doc_id;
        if (*doc_id >= target) break;
    }
    *field_mask = (uint32_t)m;
    if (r->mask == 0 || (m & r->mask)) return 1;
    return orc_reader_next(r, doc_id, freq, field_mask); /* filtered reader keeps reading (field_mask.rs) */
}

/* ------------------------------------------------------------------ leaf iterator ----------- */
typedef struct {
    OrcReader *r;
    uint32_t orig; /* index in the caller's array */
    uint64_t last_doc_id;
    uint32_t freq;
    int at_eof;
} Leaf;

static int leaf_read(Leaf *l) { /* rqe_iterators/src/inverted_index/core.rs:237-250 */
    if (l->at_eof) return 0;
    uint64_t d;
    uint32_t f, m;
    if (!orc_reader_next(l->r, &d, &f, &m)) {
        l->at_eof = 1;
        return 0;
    }
    l->last_doc_id = d;
    l->freq = f;
    return 1;
}
/* 0 = EOF, 1 = Found, 2 = NotFound (:327-350) */
static int leaf_skip_to(Leaf *l, uint64_t target) {
    if (l->at_eof) return 0;
    uint64_t d;
    uint32_t f, m;
    if (!orc_reader_seek(l->r, target, &d, &f, &m)) {
        l->at_eof = 1;
        return 0;
    }
    l->last_doc_id = d;
    l->freq = f;
    return d == target ? 1 : 2;
}
static void leaf_rewind(Leaf *l) {
    orc_reader_rewind(l->r);
    l->last_doc_id = 0;
    l->at_eof = 0;
    l->freq = 0;
}

/* ------------------------------------------------------------------ intersection ------------ */
typedef struct {
    Leaf *c;
    size_t n;
    uint64_t last_doc_id;
    int is_eof;
} Inter;

static void inter_init(Inter *it, OrcReader **children, size_t n) {
    it->c = (Leaf *)calloc(n ? n : 1, sizeof(Leaf));
    it->n = n;
    for (size_t i = 0; i < n; i++) {
        it->c[i].r = children[i];
        it->c[i].orig = (uint32_t)i;
        leaf_rewind(&it->c[i]);
    }
    /* stable sort ascending by num_estimated * weight(=1.0 for leaves), intersection.rs:110-145 */
    for (size_t i = 1; i < n; i++) {
        Leaf key = it->c[i];
        size_t j = i;
        while (j > 0 && orc_ii_num_docs(it->c[j - 1].r->ii) > orc_ii_num_docs(key.r->ii)) {
            it->c[j] = it->c[j - 1];
            j--;
        }
        it->c[j] = key;
    }
    it->last_doc_id = 0;
    it->is_eof = (n == 0);
}

/* agree_on_doc_id + find_consensus (intersection.rs:256-288). 1 = consensus on *target, 0 = EOF */
static int find_consensus(Inter *it, uint64_t *target) {
    for (;;) {
        int ahead = 0;
        for (size_t i = 0; i < it->n; i++) {
            Leaf *c = &it->c[i];
            if (c->last_doc_id == *target) continue;
            int st = leaf_skip_to(c, *target);
            if (st == 0) {
                it->is_eof = 1;
                return 0;
            }
            if (st == 2) {
                *target = c->last_doc_id;
                ahead = 1;
                break;
            }
        }
        if (!ahead) return 1;
    }
}
static void fill_hit(const Inter *it, uint64_t doc, OrcHit *h) {
    h->doc_id = doc;
    h->n_children = (uint32_t)it->n;
    for (size_t i = 0; i < it->n && i < 16; i++) {
        h->child_index[i] = it->c[i].orig;
        h->child_freq[i] = it->c[i].freq;
    }
}
static int inter_read(Inter *it, OrcHit *h) { /* :428-452 */
    if (it->is_eof) return 0;
    if (!leaf_read(&it->c[0])) {
        it->is_eof = 1;
        return 0;
    }
    uint64_t target = it->c[0].last_doc_id;
    if (!find_consensus(it, &target)) return 0;
    it->last_doc_id = target;
    if (h) fill_hit(it, target, h);
    return 1;
}
/* 0 OK(found), 1 NOTFOUND, 2 EOF (:454-506) */
static int inter_skip_to(Inter *it, uint64_t doc, OrcHit *h) {
    if (it->is_eof) return 2;
    uint64_t target = doc;
    if (!find_consensus(it, &target)) return 2;
    it->last_doc_id = target;
    if (h) fill_hit(it, target, h);
    return target == doc ? 0 : 1;
}

size_t orc_intersect(OrcReader **children, size_t n, OrcHit *hits, size_t cap) {
    Inter it;
    inter_init(&it, children, n);
    size_t m = 0;
    OrcHit h;
    while (inter_read(&it, &h)) {
        if (hits && m < cap) hits[m] = h;
        m++;
    }
    free(it.c);
    return m;
}
size_t orc_intersect_skipto(OrcReader **children, size_t n, const uint64_t *targets, size_t nt, int *status,
                            uint64_t *landed) {
    Inter it;
    inter_init(&it, children, n);
    size_t done = 0;
    for (size_t i = 0; i < nt; i++) {
        if (targets[i] <= it.last_doc_id && i > 0) { /* contract: lastDocId < target; use Read instead */
            OrcHit h;
            int ok = inter_read(&it, &h);
            status[i] = ok ? 0 : 2;
            landed[i] = it.last_doc_id;
        } else {
            status[i] = inter_skip_to(&it, targets[i], NULL);
            landed[i] = it.last_doc_id;
        }
        done++;
        if (status[i] == 2) break;
    }
    free(it.c);
    return done;
}

/* ------------------------------------------------------------------ union (flat) ------------ */
typedef struct {
    Leaf *c;
    size_t n, active;
    uint64_t last_doc_id;
    int is_eof, quick;
} Uni;

static void uni_init(Uni *u, OrcReader **children, size_t n, int quick) {
    u->c = (Leaf *)calloc(n ? n : 1, sizeof(Leaf));
    u->n = u->active = n;
    for (size_t i = 0; i < n; i++) {
        u->c[i].r = children[i];
        u->c[i].orig = (uint32_t)i;
        leaf_rewind(&u->c[i]);
    }
    u->last_doc_id = 0;
    u->is_eof = (n == 0);
    u->quick = quick;
}
static void uni_swap_remove(Uni *u, size_t idx) { /* union_flat.rs:174-180 */
    u->active--;
    if (idx < u->active) {
        Leaf t = u->c[idx];
        u->c[idx] = u->c[u->active];
        u->c[u->active] = t;
    }
}
static void uni_fill(const Uni *u, uint64_t doc, OrcHit *h) { /* build_aggregate_result :302-322 */
    h->doc_id = doc;
    h->n_children = 0;
    for (size_t i = 0; i < u->active; i++)
        if (u->c[i].last_doc_id == doc && !u->c[i].at_eof && h->n_children < 16) {
            h->child_index[h->n_children] = u->c[i].orig;
            h->child_freq[h->n_children] = u->c[i].freq;
            h->n_children++;
        }
}
static int uni_skip_to(Uni *u, uint64_t doc, OrcHit *h);

static int uni_read(Uni *u, OrcHit *h) {
    if (u->is_eof) return 0;
    if (u->quick) { /* read_quick :433-444 */
        int st = uni_skip_to(u, u->last_doc_id + 1, h);
        return st != 2;
    }
    uint64_t min_id = UINT64_MAX; /* read_full :324-348 */
    size_t i = 0;
    const uint64_t prev = u->last_doc_id;
    while (i < u->active) {
        Leaf *c = &u->c[i];
        if (prev == 0 ? (c->last_doc_id == 0) : (c->last_doc_id == prev)) {
            if ((prev == 0 && c->at_eof) || !leaf_read(c)) {
                uni_swap_remove(u, i);
                continue;
            }
        }
        if (c->last_doc_id < min_id) min_id = c->last_doc_id;
        i++;
    }
    if (min_id == UINT64_MAX) {
        u->is_eof = 1;
        return 0;
    }
    u->last_doc_id = min_id;
    if (h) uni_fill(u, min_id, h);
    return 1;
}
/* 0 OK, 1 NOTFOUND, 2 EOF */
static int uni_skip_to(Uni *u, uint64_t doc, OrcHit *h) {
    if (u->is_eof) return 2;
    uint64_t min_id = UINT64_MAX;
    size_t i = 0;
    if (u->quick) { /* skip_to_quick :446-503 */
        size_t min_idx = 0;
        while (i < u->active) {
            Leaf *c = &u->c[i];
            if (c->last_doc_id < doc) {
                int st = leaf_skip_to(c, doc);
                if (st == 1) {
                    u->last_doc_id = doc;
                    if (h) {
                        h->doc_id = doc;
                        h->n_children = 1;
                        h->child_index[0] = c->orig;
                        h->child_freq[0] = c->freq;
                    }
                    return 0;
                }
                if (st == 0) {
                    uni_swap_remove(u, i);
                    continue;
                }
                if (c->last_doc_id < min_id) {
                    min_id = c->last_doc_id;
                    min_idx = i;
                }
            } else if (c->last_doc_id == doc) {
                u->last_doc_id = doc;
                if (h) {
                    h->doc_id = doc;
                    h->n_children = 1;
                    h->child_index[0] = c->orig;
                    h->child_freq[0] = c->freq;
                }
                return 0;
            } else if (c->last_doc_id < min_id) {
                min_id = c->last_doc_id;
                min_idx = i;
            }
            i++;
        }
        if (min_id == UINT64_MAX) {
            u->is_eof = 1;
            return 2;
        }
        u->last_doc_id = min_id;
        if (h) {
            h->doc_id = min_id;
            h->n_children = 1;
            h->child_index[0] = u->c[min_idx].orig;
            h->child_freq[0] = u->c[min_idx].freq;
        }
        return 1;
    }
    while (i < u->active) { /* skip_to_full :356-424 */
        Leaf *c = &u->c[i];
        if (c->last_doc_id >= doc) {
            if (c->last_doc_id < min_id) min_id = c->last_doc_id;
            i++;
            continue;
        }
        int st = leaf_skip_to(c, doc);
        if (st == 0) {
            uni_swap_remove(u, i);
            continue;
        }
        if (c->last_doc_id < min_id) min_id = c->last_doc_id;
        i++;
    }
    if (min_id == UINT64_MAX) {
        u->is_eof = 1;
        return 2;
    }
    u->last_doc_id = min_id;
    if (h) uni_fill(u, min_id, h);
    return min_id == doc ? 0 : 1;
}

size_t orc_union(OrcReader **children, size_t n, int quick_exit, OrcHit *hits, size_t cap) {
    Uni u;
    uni_init(&u, children, n, quick_exit);
    size_t m = 0;
    OrcHit h;
    while (uni_read(&u, &h)) {
        if (hits && m < cap) hits[m] = h;
        m++;
    }
    free(u.c);
    return m;
}
size_t orc_union_skipto(OrcReader **children, size_t n, const uint64_t *targets, size_t nt, int *status,
                        uint64_t *landed) {
    Uni u;
    uni_init(&u, children, n, 0);
    size_t done = 0;
    for (size_t i = 0; i < nt; i++) {
        status[i] = uni_skip_to(&u, targets[i], NULL);
        landed[i] = u.last_doc_id;
        done++;
        if (status[i] == 2) break;
    }
    free(u.c);
    return done;
}

/* ------------------------------------------------------------------ idf --------------------- */
double orc_idf(uint64_t total_docs, uint64_t term_docs) { /* idf/src/lib.rs:36-70 */
    if (term_docs == 0) term_docs = 1;
    double value = 1.0 + (double)(total_docs + 1) / (double)term_docs;
    uint64_t bits;
    memcpy(&bits, &value, 8);
    return (double)((int)((bits >> 52) & 0x7FF) - 1023);
}
double orc_idf_bm25(uint64_t total_docs, uint64_t term_docs) { /* :103-110 */
    if (total_docs < term_docs) total_docs = term_docs;
    double total = (double)total_docs, term = (double)term_docs;
    return log(1.0 + (total - term + 0.5) / (term + 0.5));
}

/* ------------------------------------------------------------------ synthetic postings ------ */
/* Zipf vocabulary, hashed membership, tf = 1 + min(254, Geom(0.5)), docLen 50..500. */
uint64_t orc_synth_df(uint64_t n_docs, uint64_t rank) {
    uint64_t df = (uint64_t)((double)n_docs * 0.2 / (double)rank);
    return df > n_docs ? n_docs : df;
}
int orc_synth_member(uint64_t n_docs, uint64_t rank, uint64_t doc, uint32_t *tf) {
    const uint64_t h = orc_mix64(7, rank, doc);
    /* P(member) = df/N via a 32-bit threshold (exact integer arithmetic, same on device) */
    const uint64_t thresh = (orc_synth_df(n_docs, rank) << 32) / n_docs;
    if ((h >> 32) >= thresh) return 0;
    if (tf) {
        uint32_t low = (uint32_t)h;
        uint32_t g = 0;
        while (g < 31 && (low & 1u)) { /* Geom(0.5): number of trailing ones */
            g++;
            low >>= 1;
        }
        *tf = 1 + g;
    }
    return 1;
}
uint32_t orc_synth_doclen(uint64_t doc) { return 50u + (uint32_t)(orc_mix64(11, doc, 0) % 451u); }

/* ------------------------------------------------------------------ bulk fill + CPU baseline -- */
size_t orc_ii_fill_synth(OrcInvIndex *ii, uint64_t n_docs, uint64_t rank) {
    size_t added = 0;
    for (uint64_t d = 1; d <= n_docs; d++) {
        uint32_t tf;
        if (orc_synth_member(n_docs, rank, d, &tf)) {
            orc_ii_add(ii, d, tf, 1, NULL, 0);
            added++;
        }
    }
    return added;
}

#include <pthread.h>
#include <time.h>
double orc_score(int scorer, const OrcIndexStats *st, const OrcScoreDoc *d, int slop, double min_score, double tanh_factor);

typedef struct {
    OrcInvIndex **terms; /* nq * 3 indexes */
    const uint32_t *doc_len;
    uint64_t n_docs;
    double avg_doc_len;
    size_t nq, top_n;
    size_t *next;
    pthread_mutex_t *mu;
    uint64_t *out_ids;   /* nq * top_n */
    double *out_scores;
    size_t *out_hits;
} SearchWork;

/* One FT.SEARCH "t1 t2 t3" worth of hot path on the CPU: readers over the encoded blocks ->
 * Intersection::read loop -> BM25STD per hit -> top-N by (score desc, docId asc)
 * (src/result_processor.c:317-379, 570-603, 752-850). */
static void search_one(SearchWork *w, size_t q) {
    OrcReader *r[3];
    for (int i = 0; i < 3; i++) r[i] = orc_reader_new(w->terms[q * 3 + i], 0);
    Inter it;
    inter_init(&it, r, 3);
    double idf[3], w1[3] = {1.0, 1.0, 1.0}, zero[3] = {0, 0, 0};
    for (size_t i = 0; i < 3; i++) idf[i] = orc_idf_bm25(w->n_docs, orc_ii_num_docs(it.c[i].r->ii));
    OrcIndexStats st = {w->n_docs, 0, w->avg_doc_len};
    uint64_t *ids = w->out_ids + q * w->top_n;
    double *sc = w->out_scores + q * w->top_n;
    size_t have = 0, hits = 0;
    OrcHit h;
    while (inter_read(&it, &h)) {
        uint32_t fr[3] = {h.child_freq[0], h.child_freq[1], h.child_freq[2]};
        OrcScoreDoc d = {3, fr, zero, idf, w1, 1.0, w->doc_len[h.doc_id], 1, 1.0f};
        double s = orc_score(ORC_SCORER_BM25STD, &st, &d, 1, 0, 0);
        hits++;
        /* insertion into a small sorted array = the RPSorter heap for tiny N */
        size_t pos = have;
        while (pos > 0 && (sc[pos - 1] < s)) pos--;
        if (pos < w->top_n) {
            size_t end = have < w->top_n ? have : w->top_n - 1;
            for (size_t k = end; k > pos; k--) {
                sc[k] = sc[k - 1];
                ids[k] = ids[k - 1];
            }
            sc[pos] = s;
            ids[pos] = h.doc_id;
            if (have < w->top_n) have++;
        }
    }
    for (size_t k = have; k < w->top_n; k++) {
        ids[k] = 0;
        sc[k] = 0;
    }
    w->out_hits[q] = hits;
    free(it.c);
    for (int i = 0; i < 3; i++) orc_reader_free(r[i]);
}
static void *search_worker(void *arg) {
    SearchWork *w = (SearchWork *)arg;
    for (;;) {
        pthread_mutex_lock(w->mu);
        size_t q = (*w->next)++;
        pthread_mutex_unlock(w->mu);
        if (q >= w->nq) break;
        search_one(w, q);
    }
    return NULL;
}
/* Wall seconds for nq 3-term AND + BM25STD + top-N queries over `terms` (nq*3 indexes) on nthreads
 * threads (one query per thread, like the reference's worker pool). */
double orc_time_search3(OrcInvIndex **terms, size_t nq, const uint32_t *doc_len, uint64_t n_docs, double avg_doc_len,
                        size_t top_n, int nthreads, uint64_t *out_ids, double *out_scores, size_t *out_hits) {
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    size_t next = 0;
    SearchWork w = {terms, doc_len, n_docs, avg_doc_len, nq, top_n, &next, &mu, out_ids, out_scores, out_hits};
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, search_worker, &w);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ proximity (slop / in-order) ----------
 * RS/index_result/src/core/proximity.rs: OffsetIter::Term (:45-52: varint deltas accumulated into positions),
 * within_range_in_order (:127-180), within_range_unordered (:184-220), is_within_range (:262-299: children without offsets
 * are left out, <= 1 stream left is trivially in range).  Term children only (no Merge streams). */
typedef struct {
    const uint8_t *p, *end;
    uint32_t last;
} OffCursor;
static int off_next(OffCursor *c, uint32_t *pos) {
    if (c->p >= c->end) return 0;
    uint64_t d = 0;
    /* a truncated varint is an error in the reference (read(...).ok()? -> None): treat as EOF */
    const uint8_t *q = c->p;
    uint8_t b = *q++;
    uint64_t val = b & 0x7f;
    while (b & 0x80) {
        if (q >= c->end) return 0;
        val += 1;
        b = *q++;
        val = (val << 7) | (b & 0x7f);
    }
    d = val;
    c->p = q;
    c->last = c->last + (uint32_t)d; /* wrapping_add */
    *pos = c->last;
    return 1;
}
int orc_within_range(size_t n_children, const uint8_t *const *offsets, const size_t *lens, int has_slop, uint32_t max_slop_in,
                     int in_order) {
    OffCursor it[64];
    size_t n = 0;
    for (size_t i = 0; i < n_children && n < 64; i++)
        if (lens[i]) { /* has_offsets */
            it[n].p = offsets[i];
            it[n].end = offsets[i] + lens[i];
            it[n].last = 0;
            n++;
        }
    if (n_children <= 1 || n <= 1) return 1;
    const uint32_t max_slop = has_slop ? max_slop_in : 0xFFFFFFFFu;
    uint32_t positions[64];
    if (in_order) {
        for (size_t i = 0; i < n; i++) positions[i] = 0;
        for (;;) {
            int32_t span = 0;
            int over = 0;
            for (size_t i = 0; i < n; i++) {
                uint32_t pos;
                if (i == 0) {
                    if (!off_next(&it[0], &pos)) return 0;
                } else {
                    pos = positions[i];
                }
                const uint32_t last_pos = i == 0 ? 0u : positions[i - 1];
                while (pos < last_pos)
                    if (!off_next(&it[i], &pos)) return 0;
                positions[i] = pos;
                if (i > 0) {
                    span += (int32_t)pos - (int32_t)last_pos - 1;
                    if (span > 0 && (uint32_t)span > max_slop) {
                        over = 1;
                        break;
                    }
                }
            }
            if (!over) return 1;
        }
    }
    for (size_t i = 0; i < n; i++)
        if (!off_next(&it[i], &positions[i])) return 0;
    uint32_t max_pos = 0;
    for (size_t i = 0; i < n; i++)
        if (positions[i] >= max_pos) max_pos = positions[i];
    for (;;) {
        uint32_t min_pos = 0xFFFFFFFFu;
        size_t min_idx = 0;
        for (size_t i = 0; i < n; i++)
            if (positions[i] < min_pos) { /* min_by_key: the first minimum */
                min_pos = positions[i];
                min_idx = i;
            }
        if (min_pos != max_pos) {
            const int32_t span = (int32_t)max_pos - (int32_t)min_pos - ((int32_t)n - 1);
            if (span < 0 || (uint32_t)span <= max_slop) return 1;
        }
        uint32_t np;
        if (!off_next(&it[min_idx], &np)) break;
        positions[min_idx] = np;
        if (np > max_pos) max_pos = np;
    }
    return 0;
}

/* ------------------------------------------------------------------ numeric codec ----------- */
/* RS/inverted_index/src/codec/numeric.rs: header byte = (type-specific << 5) | (type << 3) | delta_bytes, then the docId delta
 * (0-7 bytes LE, trailing zero bytes trimmed), then the value.  Value::from (:790-838) picks the representation; encode_value
 * (:363-520) writes it; decode (:553-626) reads it back.  compress_floats = NumericFloatCompression (f32 when the loss is below
 * 0.01).  Golden bytes: inverted_index/tests/integration/codec/numeric.rs:220-600. */
static size_t trim_le(uint64_t v, uint8_t *out) { /* little endian without trailing zero bytes; returns the length (0 for 0) */
    size_t n = 0;
    while (v) {
        out[n++] = (uint8_t)v;
        v >>= 8;
    }
    return n;
}
size_t orc_numeric_encode(uint64_t delta, double value, int compress_floats, uint8_t *out) {
    uint8_t dbytes[8], vbytes[8];
    const size_t nd = trim_le(delta, dbytes); /* callers keep delta below 2^56 (NumericDelta::from_u64 :297-305) */
    size_t nv = 0;
    uint8_t type = 0, upper = 0;
    const double abs_val = fabs(value);
    /* `abs_val as u64` saturates in Rust: NaN -> 0, >= 2^64 -> u64::MAX */
    uint64_t u64_val;
    if (!(abs_val == abs_val))
        u64_val = 0;
    else if (abs_val >= 18446744073709551616.0)
        u64_val = UINT64_MAX;
    else
        u64_val = (uint64_t)abs_val;
    if ((double)u64_val == abs_val) {
        const uint64_t tiny = u64_val & 7;
        if ((double)tiny == value) { /* TINY: 0..7 (and -0.0, which compares equal to 0) */
            type = 0;
            upper = (uint8_t)u64_val;
        } else {
            type = signbit(value) ? 3 : 2; /* INT_NEG / INT_POS: magnitude, trailing zeros trimmed, length - 1 in the header */
            nv = trim_le(u64_val, vbytes);
            upper = (uint8_t)(nv - 1);
        }
    } else if (isinf(value)) {
        type = 1;
        upper = value > 0 ? 1 : 3; /* FLOAT_INFINITE / FLOAT_NEGATIVE_INFINITE */
    } else {
        type = 1;
        const float f32v = (float)abs_val;
        const double back = (double)f32v;
        if (back == abs_val || (compress_floats && fabs(abs_val - (double)f32v) < 0.01)) {
            if (f32v == 0.0f) { /* collapsed onto zero: the canonical zero */
                type = 0;
                upper = 0;
            } else {
                upper = signbit(value) ? 2 : 0; /* FLOAT32_NEGATIVE / FLOAT32_POSITIVE: the magnitude's f32 bits */
                memcpy(vbytes, &f32v, 4);
                nv = 4;
            }
        } else {
            upper = signbit(value) ? 6 : 4; /* FLOAT64_NEGATIVE / FLOAT64_POSITIVE */
            memcpy(vbytes, &abs_val, 8);
            nv = 8;
        }
    }
    out[0] = (uint8_t)((upper << 5) | (type << 3) | (uint8_t)nd);
    memcpy(out + 1, dbytes, nd);
    memcpy(out + 1 + nd, vbytes, nv);
    return 1 + nd + nv;
}
size_t orc_numeric_decode(const uint8_t *in, uint64_t *delta, double *value) {
    const uint8_t header = in[0];
    const size_t nd = header & 7;
    const uint8_t type = (header >> 3) & 3, upper = header >> 5;
    uint64_t d = 0;
    for (size_t i = 0; i < nd; i++) d |= (uint64_t)in[1 + i] << (8 * i);
    *delta = d;
    const uint8_t *v = in + 1 + nd;
    size_t nv = 0;
    if (type == 0) {
        *value = (double)upper;
    } else if (type == 2 || type == 3) {
        nv = (size_t)upper + 1;
        uint64_t m = 0;
        for (size_t i = 0; i < nv; i++) m |= (uint64_t)v[i] << (8 * i);
        *value = type == 3 ? copysign((double)m, -1.0) : (double)m;
    } else if (upper == 0 || upper == 2) {
        float f;
        memcpy(&f, v, 4);
        nv = 4;
        *value = upper == 2 ? (double)copysignf(f, -1.0f) : (double)f;
    } else if (upper == 4 || upper == 6) {
        double f;
        memcpy(&f, v, 8);
        nv = 8;
        *value = upper == 6 ? copysign(f, -1.0) : f;
    } else {
        *value = (upper == 1 || upper == 5) ? INFINITY : -INFINITY; /* 0b101 / 0b111 are the unused twins (:607-616) */
    }
    return 1 + nd + nv;
}
/* NumericFilter::value_in_range, RS/inverted_index/src/reader/numeric.rs:80-85 */
int orc_numeric_in_range(double value, double min, double max, int min_inclusive, int max_inclusive) {
    const int min_ok = value > min || (min_inclusive && value == min);
    const int max_ok = value < max || (max_inclusive && value == max);
    return min_ok && max_ok;
}
