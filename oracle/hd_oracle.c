/*
 * hd_oracle.c -- CPU restatement of the reference decode path (see hd_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into the product library.
 * Citations are relative to /root/reference.
 */
#include "hd_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- bit access: unit 0 bit 31 is stream bit 0 ---------------------------
 * gpuhd/encoder/src/llhuffman_encoder.cc:216-231 (window <<= len; += code) */
static inline uint32_t unit_at(const uint32_t *units, uint64_t n_units,
                               uint64_t i) {
    return i < n_units ? units[i] : 0u;
}

/* next k (<=32) bits starting at absolute bit position p, right-aligned */
static inline uint32_t peek_bits(const uint32_t *units, uint64_t n_units,
                                 uint64_t p, uint32_t k) {
    uint64_t w = p >> 5;
    uint32_t off = (uint32_t)(p & 31);
    uint64_t two = ((uint64_t)unit_at(units, n_units, w) << 32) |
                   unit_at(units, n_units, w + 1);
    return (uint32_t)((two << off) >> (64 - k));
}

int oracle_canonical_codes(const uint8_t *syms, const uint8_t *lengths, int n,
                           uint32_t codes[256], uint8_t lens[256]) {
    memset(codes, 0, 256 * sizeof(uint32_t));
    memset(lens, 0, 256);
    if (n <= 0 || n > 256) return -1;
    uint32_t code = 0;
    uint32_t cur = lengths[0];
    for (int i = 0; i < n; ++i) {
        if (lengths[i] != cur || cur == 0 || cur > 32) return -1;
        codes[syms[i]] = code;
        lens[syms[i]] = (uint8_t)cur;
        uint32_t next = (i + 1 < n) ? lengths[i + 1] : cur;
        if (next < cur) return -1;
        code = (code + 1) << (next - cur);
        cur = next;
    }
    return 0;
}

void oracle_build_flat_lut(const uint32_t codes[256], const uint8_t lens[256],
                           uint32_t lut_bits, oracle_lut_entry *lut) {
    memset(lut, 0, ((size_t)1 << lut_bits) * sizeof(oracle_lut_entry));
    for (int s = 0; s < 256; ++s) {
        uint32_t len = lens[s];
        if (len == 0 || len > lut_bits) continue;
        uint32_t shift = lut_bits - len;
        uint32_t base = codes[s] << shift;
        for (uint32_t j = 0; j < (1u << shift); ++j) {
            lut[base + j].num_bits = (uint8_t)len;
            lut[base + j].symbol = (uint8_t)s;
        }
    }
}

uint64_t oracle_pack(const uint8_t *in, uint64_t n, const uint32_t codes[256],
                     const uint8_t lens[256], uint32_t *units_out,
                     uint64_t cap_units, int flush_tail, uint8_t *gap_nibbles,
                     uint64_t cap_segments) {
    memset(units_out, 0, cap_units * sizeof(uint32_t));
    if (gap_nibbles) memset(gap_nibbles, 0, cap_segments);
    uint64_t p = 0; /* absolute bit position */
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t len = lens[in[i]];
        uint32_t code = codes[in[i]];
        uint64_t e = p + len;
        if (gap_nibbles && (e >> 7) != (p >> 7) && (p >> 7) < cap_segments)
            gap_nibbles[p >> 7] = (uint8_t)(e & 15);
        /* write code bits MSB-first */
        uint64_t w = p >> 5;
        uint32_t off = (uint32_t)(p & 31);
        if (off + len <= 32) {
            if (w < cap_units) units_out[w] |= code << (32 - off - len);
        } else {
            uint32_t spill = off + len - 32;
            if (w < cap_units) units_out[w] |= code >> spill;
            int last = (i + 1 == n);
            /* D1: the reference llhuff packer never flushes the remainder of
             * the final codeword (llhuffman_encoder.cc:213-237) */
            if ((flush_tail || !last) && w + 1 < cap_units)
                units_out[w + 1] |= code << (32 - spill);
        }
        p = e;
    }
    return p;
}

void oracle_pack_gap_words(const uint8_t *gap_nibbles, uint64_t n_segments,
                           uint32_t *words_out) {
    uint64_t n_words = (n_segments + 7) / 8;
    memset(words_out, 0, n_words * sizeof(uint32_t));
    for (uint64_t i = 0; i < n_segments; ++i)
        words_out[i >> 3] |= (uint32_t)(gap_nibbles[i] & 15) << (4 * (i & 7));
}

uint64_t oracle_decode_flat(const uint32_t *units, uint64_t n_units,
                            const oracle_lut_entry *lut, uint32_t lut_bits,
                            uint8_t *out, uint64_t n_symbols) {
    uint64_t p = 0;
    uint64_t total_bits = n_units * 32;
    uint64_t k = 0;
    while (k < n_symbols && p < total_bits + 32) {
        oracle_lut_entry hit = lut[peek_bits(units, n_units, p, lut_bits)];
        if (hit.num_bits == 0) break; /* incomplete code hit in pad bits */
        out[k++] = hit.symbol;
        p += hit.num_bits;
    }
    return k;
}

uint64_t oracle_sync_points(const uint32_t *units, uint64_t n_units,
                            const oracle_lut_entry *lut, uint32_t lut_bits,
                            uint64_t n_symbols, uint32_t subseq_bits,
                            uint8_t *start_off, uint32_t *counts,
                            uint64_t cap_subseq) {
    uint64_t total_bits = n_units * 32;
    uint64_t n_sub = (total_bits + subseq_bits - 1) / subseq_bits;
    if (n_sub > cap_subseq) n_sub = cap_subseq;
    for (uint64_t s = 0; s < n_sub; ++s) {
        start_off[s] = 0;
        counts[s] = 0;
    }
    uint64_t p = 0, k = 0;
    uint64_t cur = 0; /* subsequence whose start offset is known */
    while (k < n_symbols) {
        uint64_t s = p / subseq_bits;
        if (s >= n_sub) break;
        while (cur < s) { /* first codeword at/after the start of cur+1.. */
            ++cur;
            /* only the subsequence actually containing p gets p's offset; a
             * skipped one (impossible when len <= subseq_bits) would keep 0 */
            if (cur == s) start_off[cur] = (uint8_t)(p - s * subseq_bits);
        }
        oracle_lut_entry hit = lut[peek_bits(units, n_units, p, lut_bits)];
        if (hit.num_bits == 0) break;
        counts[s] += 1;
        ++k;
        p += hit.num_bits;
    }
    /* the parse position after the last symbol defines the start offset of the
     * following subsequence when it lands exactly there */
    {
        uint64_t s = p / subseq_bits;
        if (s < n_sub && s > cur) start_off[s] = (uint8_t)(p - s * subseq_bits);
    }
    return n_sub;
}

uint64_t oracle_decode_gap(const uint32_t *units, uint64_t n_units,
                           const uint32_t *gap_words, uint64_t n_segments,
                           const oracle_lut_entry *lut, uint32_t lut_bits,
                           uint8_t *out, uint64_t n_symbols) {
    uint64_t k = 0;
    for (uint64_t seg = 0; seg < n_segments && k < n_symbols; ++seg) {
        uint32_t gap = 0;
        if (seg > 0) {
            uint64_t j = seg - 1;
            gap = (gap_words[j >> 3] >> (4 * (j & 7))) & 15u;
        }
        uint64_t p = seg * 128 + gap;
        uint64_t end = (seg + 1) * 128;
        while (p < end && k < n_symbols) {
            oracle_lut_entry hit = lut[peek_bits(units, n_units, p, lut_bits)];
            if (hit.num_bits == 0) { p = end; break; }
            out[k++] = hit.symbol;
            p += hit.num_bits;
        }
    }
    return k;
}

/* ---- .huff container ----------------------------------------------------- */

static uint32_t rd_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
           ((uint32_t)p[3] << 24);
}
static void wr_u32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

int oracle_huff_parse(const void *file, size_t size, oracle_huff_view *v) {
    const uint8_t *p = (const uint8_t *)file;
    if (size < 8) return -1;
    uint64_t sc = (uint64_t)rd_u32(p) | ((uint64_t)rd_u32(p + 4) << 32);
    if (sc == 0 || sc > 256) return -2;
    size_t off = 8;
    if (size < off + 2 * sc + 12) return -3;
    v->symbol_count = sc;
    v->sym_len_pairs = p + off;
    off += 2 * sc;
    v->original_size = rd_u32(p + off);
    v->n_units = rd_u32(p + off + 4);
    v->n_segments = rd_u32(p + off + 8);
    off += 12;
    size_t gap_words = ((size_t)v->n_segments + 7) / 8;
    if (size < off + 4 * (gap_words + (size_t)v->n_units)) return -4;
    /* header length 8+2*sc+12 is a multiple of 4 only when sc is even; the
     * payload may be unaligned, callers copy before use */
    v->gap_words = (const uint32_t *)(p + off);
    v->units = (const uint32_t *)(p + off + 4 * gap_words);
    return 0;
}

size_t oracle_huff_size(uint64_t symbol_count, uint32_t n_units,
                        uint32_t n_segments) {
    return 8 + 2 * (size_t)symbol_count + 12 +
           4 * (((size_t)n_segments + 7) / 8 + (size_t)n_units);
}

size_t oracle_huff_write(void *dst, const uint8_t *syms, const uint8_t *lengths,
                         uint64_t symbol_count, uint32_t original_size,
                         const uint32_t *units, uint32_t n_units,
                         const uint32_t *gap_words, uint32_t n_segments) {
    uint8_t *p = (uint8_t *)dst;
    wr_u32(p, (uint32_t)symbol_count);
    wr_u32(p + 4, (uint32_t)(symbol_count >> 32));
    size_t off = 8;
    for (uint64_t i = 0; i < symbol_count; ++i) {
        p[off++] = syms[i];
        p[off++] = lengths[i];
    }
    wr_u32(p + off, original_size);
    wr_u32(p + off + 4, n_units);
    wr_u32(p + off + 8, n_segments);
    off += 12;
    size_t gw = ((size_t)n_segments + 7) / 8;
    for (size_t i = 0; i < gw; ++i, off += 4) wr_u32(p + off, gap_words[i]);
    for (size_t i = 0; i < n_units; ++i, off += 4) wr_u32(p + off, units[i]);
    return off;
}

int64_t oracle_huff_decode(const void *file, size_t size, uint8_t *out,
                           uint64_t out_cap) {
    oracle_huff_view v;
    int rc = oracle_huff_parse(file, size, &v);
    if (rc) return rc;
    uint8_t syms[256], lengths[256];
    uint32_t maxlen = 0;
    for (uint64_t i = 0; i < v.symbol_count; ++i) {
        syms[i] = v.sym_len_pairs[2 * i];
        lengths[i] = v.sym_len_pairs[2 * i + 1];
        if (lengths[i] > maxlen) maxlen = lengths[i];
    }
    if (maxlen == 0 || maxlen > 16) return -5;
    uint32_t codes[256];
    uint8_t lens[256];
    if (oracle_canonical_codes(syms, lengths, (int)v.symbol_count, codes, lens))
        return -6;
    size_t gw = ((size_t)v.n_segments + 7) / 8;
    uint32_t *buf = (uint32_t *)malloc(4 * (gw + (size_t)v.n_units + 1));
    oracle_lut_entry *lut =
        (oracle_lut_entry *)malloc(sizeof(oracle_lut_entry) << maxlen);
    if (!buf || !lut) { free(buf); free(lut); return -7; }
    memcpy(buf, v.gap_words, 4 * gw);
    memcpy(buf + gw, v.units, 4 * (size_t)v.n_units);
    oracle_build_flat_lut(codes, lens, maxlen, lut);
    uint64_t want = v.original_size < out_cap ? v.original_size : out_cap;
    uint64_t got = oracle_decode_gap(buf + gw, v.n_units, buf, v.n_segments, lut,
                                     maxlen, out, want);
    free(buf);
    free(lut);
    return (int64_t)got;
}
