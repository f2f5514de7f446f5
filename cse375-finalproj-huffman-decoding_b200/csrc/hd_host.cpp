// hd_host.cpp -- host side of the B200 Huffman decoder: code-table handling,
// the .huff container, and the input-side producers (length-limited code
// construction, multi-threaded bit packer with gap-array recording, synthetic
// data).  No CUDA in this file.  Reference citations are relative to the
// reference repository root.
#include "hd_internal.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <random>
#include <thread>

namespace hd {

// --------------------------------------------------------------- code tables

static void finish_code(HostCode &c) {
    c.n_present = 0;
    c.min_len = 0;
    c.max_len = 0;
    std::vector<int> syms;
    for (int s = 0; s < 256; ++s) {
        if (!c.len[s]) continue;
        syms.push_back(s);
        c.n_present++;
        c.max_len = std::max<uint32_t>(c.max_len, c.len[s]);
        c.min_len = c.min_len ? std::min<uint32_t>(c.min_len, c.len[s]) : c.len[s];
    }
    std::sort(syms.begin(), syms.end(), [&](int a, int b) {
        if (c.len[a] != c.len[b]) return c.len[a] < c.len[b];
        return c.code[a] < c.code[b];
    });
    std::memset(c.order, 0, sizeof(c.order));
    for (size_t i = 0; i < syms.size(); ++i) c.order[i] = (uint8_t)syms[i];
}

// Recover {code,len} per symbol from the reference's flat LUT
// (gpuhd/encoder/src/llhuffman_encoder.cc:240-262: entry (code<<shift)+j).
bool code_from_flat_lut(const hd_lut_entry *lut, uint32_t lut_bits, HostCode &c) {
    if (!lut || lut_bits == 0 || lut_bits > kMaxCodeLen) return false;
    std::memset(c.code, 0, sizeof(c.code));
    std::memset(c.len, 0, sizeof(c.len));
    const uint32_t n = 1u << lut_bits;
    for (uint32_t i = 0; i < n;) {
        uint32_t len = lut[i].num_bits;
        if (len == 0) { ++i; continue; }      // hole of an incomplete code
        if (len > lut_bits) return false;
        uint32_t span = 1u << (lut_bits - len);
        if (i & (span - 1)) return false;     // run must be aligned
        uint8_t sym = lut[i].symbol;
        for (uint32_t j = 0; j < span; ++j)
            if (lut[i + j].num_bits != len || lut[i + j].symbol != sym) return false;
        if (c.len[sym]) return false;         // symbol with two codes
        c.len[sym] = (uint8_t)len;
        c.code[sym] = i >> (lut_bits - len);
        i += span;
    }
    finish_code(c);
    return c.n_present > 0;
}

// Canonical assignment in list order (decoder/src/get_table.cpp:17-20).
bool code_from_pairs(const hd_sym_len *pairs, uint32_t n, HostCode &c) {
    if (!pairs || n == 0 || n > 256) return false;
    std::memset(c.code, 0, sizeof(c.code));
    std::memset(c.len, 0, sizeof(c.len));
    uint64_t code = 0;
    uint32_t cur = pairs[0].length;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t len = pairs[i].length;
        if (len == 0 || len > kMaxCodeLen || len != cur) return false;
        if (c.len[pairs[i].symbol]) return false;
        if (code >> len) return false;        // Kraft overflow
        c.code[pairs[i].symbol] = (uint32_t)code;
        c.len[pairs[i].symbol] = (uint8_t)len;
        uint32_t next = (i + 1 < n) ? pairs[i + 1].length : len;
        if (next < len) return false;
        code = (code + 1) << (next - len);
        cur = next;
    }
    finish_code(c);
    return true;
}

// Build every device table from a host code (see TableBlob in hd_internal.h).
void build_table_blob(const HostCode &c, TableBlob &blob) {
    const uint32_t maxl = c.max_len;
    uint32_t lm = std::min<uint32_t>(kMaxLutBits, std::max<uint32_t>(maxl, std::min<uint32_t>(kMaxLutBits, 2 * maxl)));
    if (lm < 1) lm = 1;
    blob.lut_bits = lm;
    blob.min_len = c.min_len;
    blob.max_len = maxl;
    // The resolver (LONG kernel variants) is needed for codes longer than the primary LUT, and for INCOMPLETE codes:
    // their LUT has holes (windows that match no codeword), which the fast variants' symbol merge does not handle.
    uint64_t kraft = 0;
    for (int s = 0; s < 256; ++s)
        if (c.len[s]) kraft += 1ull << (kMaxCodeLen - c.len[s]);
    blob.has_long = (maxl > lm || kraft != (1ull << kMaxCodeLen)) ? 1u : 0u;

    // full-width helper table on the host: index = next maxl bits
    std::vector<uint16_t> flat((size_t)1 << maxl, 0);   // len<<8 | sym, 0 = hole
    for (int s = 0; s < 256; ++s) {
        uint32_t len = c.len[s];
        if (!len) continue;
        uint32_t span = 1u << (maxl - len);
        uint32_t base = c.code[s] << (maxl - len);
        for (uint32_t j = 0; j < span; ++j) flat[base + j] = (uint16_t)((len << 8) | s);
    }
    auto first_code = [&](uint32_t window, uint32_t pos, uint32_t &len, uint32_t &sym) -> bool {
        // window: lm bits, right aligned; look at bits [pos, lm)
        uint32_t avail = lm - pos;
        uint32_t rest = window & ((1u << avail) - 1u);   // avail bits
        uint32_t idx = (avail >= maxl) ? (rest >> (avail - maxl)) : (rest << (maxl - avail));
        uint16_t e = flat[idx];
        len = e >> 8;
        sym = e & 0xFF;
        return len != 0 && len <= avail;
    };

    const uint32_t n = 1u << lm;
    blob.off_sync = n * 4;
    blob.off_len = blob.off_sync + n * 2;
    blob.off_order = blob.off_len + 256;
    blob.off_long = blob.off_order + 256;
    blob.total_bytes = blob.off_long + (uint32_t)sizeof(LongTables);
    blob.total_bytes = (blob.total_bytes + 15u) & ~15u;
    blob.bytes.assign(blob.total_bytes, 0);
    uint32_t *dec = reinterpret_cast<uint32_t *>(blob.bytes.data());
    uint16_t *syn = reinterpret_cast<uint16_t *>(blob.bytes.data() + blob.off_sync);

    for (uint32_t w = 0; w < n; ++w) {
        uint32_t len0, sym0;
        if (!first_code(w, 0, len0, sym0)) {
            dec[w] = 0;                // (has_long is set) -> resolve_long finds the long codeword or skips 1 bit
            syn[w] = 0;
            continue;
        }
        uint32_t bits = len0, count = 1;
        uint32_t len1, sym1 = 0;
        if (first_code(w, bits, len1, sym1)) {
            bits += len1;
            count = 2;
        } else {
            sym1 = 0;
        }
        dec[w] = lut_entry(bits, sym0, sym1, count);
        // sync: jump over every complete code in the window
        uint32_t jump = len0, l, s;
        while (jump < lm && first_code(w, jump, l, s)) jump += l;
        syn[w] = (uint16_t)(jump | (len0 << 8));
    }
    std::memcpy(blob.bytes.data() + blob.off_len, c.len, 256);
    std::memcpy(blob.bytes.data() + blob.off_order, c.order, 256);
    LongTables lt;
    std::memset(&lt, 0, sizeof(lt));
    {
        uint32_t idx = 0;
        for (uint32_t i = 0; i < c.n_present;) {
            uint32_t s = c.order[i];
            uint32_t len = c.len[s];
            lt.first[len] = c.code[s];
            lt.offset[len] = idx;
            uint32_t k = 0;
            while (i < c.n_present && c.len[c.order[i]] == len) { ++i; ++k; }
            lt.count[len] = k;
            idx += k;
        }
    }
    std::memcpy(blob.bytes.data() + blob.off_long, &lt, sizeof(lt));
}

// ------------------------------------------------------- length-limited code
// Package-merge over integer weights.  Same optimisation problem as
// gpuhd/encoder/src/llhuffman_encoder.cc:17-160 (limit 11, fp32 weights) and
// Huffman_coding_Gap_arrays/encoder/src/package_merge.cpp:107-166 (limit 16).
static bool package_merge(const std::vector<uint64_t> &w_sorted, uint32_t max_len,
                          std::vector<uint32_t> &len_out) {
    const size_t n = w_sorted.size();
    len_out.assign(n, 0);
    if (n == 1) { len_out[0] = 1; return true; }
    if (((size_t)1 << max_len) < n) return false;
    struct Node { uint64_t w; bool pkg; };
    std::vector<std::vector<Node>> lists(max_len);
    for (uint32_t l = 0; l < max_len; ++l) {
        std::vector<Node> pk;
        if (l > 0) {
            const auto &prev = lists[l - 1];
            for (size_t j = 0; j + 1 < prev.size(); j += 2)
                pk.push_back({prev[j].w + prev[j + 1].w, true});
        }
        auto &cur = lists[l];
        cur.reserve(pk.size() + n);
        size_t a = 0, b = 0;
        while (a < n || b < pk.size()) {
            // items first on ties: keeps packages (deeper codes) from displacing leaves
            if (b >= pk.size() || (a < n && w_sorted[a] <= pk[b].w)) cur.push_back({w_sorted[a++], false});
            else cur.push_back(pk[b++]);
        }
    }
    size_t take = 2 * n - 2;
    for (int l = (int)max_len - 1; l >= 0 && take > 0; --l) {
        const auto &cur = lists[l];
        if (take > cur.size()) take = cur.size();
        size_t items = 0, pkgs = 0;
        for (size_t j = 0; j < take; ++j) (cur[j].pkg ? pkgs : items)++;
        for (size_t j = 0; j < items; ++j) len_out[j]++;
        take = 2 * pkgs;
    }
    return true;
}

}  // namespace hd

using namespace hd;

extern "C" {

hd_status hd_build_code(const uint8_t *data, uint64_t n, uint32_t max_len,
                        hd_sym_len *pairs_out, uint32_t *n_pairs_out) {
    if (!data || !pairs_out || !n_pairs_out || n == 0 || max_len == 0 || max_len > kMaxCodeLen)
        return HD_ERR_INVALID;
    // histogram, parallel over hardware threads
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < (1u << 20)) nt = 1;
    std::vector<std::vector<uint64_t>> part(nt, std::vector<uint64_t>(256, 0));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            uint64_t b = n * t / nt, e = n * (t + 1) / nt;
            uint64_t h[4][256] = {};
            uint64_t i = b;
            for (; i + 4 <= e; i += 4) { h[0][data[i]]++; h[1][data[i + 1]]++; h[2][data[i + 2]]++; h[3][data[i + 3]]++; }
            for (; i < e; ++i) h[0][data[i]]++;
            for (int s = 0; s < 256; ++s) part[t][s] = h[0][s] + h[1][s] + h[2][s] + h[3][s];
        });
    }
    for (auto &x : th) x.join();
    uint64_t hist[256] = {};
    for (unsigned t = 0; t < nt; ++t) for (int s = 0; s < 256; ++s) hist[s] += part[t][s];
    return hd_build_code_from_histogram(hist, max_len, pairs_out, n_pairs_out);
}

hd_status hd_build_code_from_histogram(const uint64_t *hist, uint32_t max_len, hd_sym_len *pairs_out,
                                       uint32_t *n_pairs_out) {
    if (!hist || !pairs_out || !n_pairs_out || max_len == 0 || max_len > kMaxCodeLen) return HD_ERR_INVALID;
    std::vector<int> syms;
    for (int s = 0; s < 256; ++s) if (hist[s]) syms.push_back(s);
    if (syms.empty() || ((size_t)1 << max_len) < syms.size()) return HD_ERR_INVALID;
    std::sort(syms.begin(), syms.end(), [&](int a, int b) {
        if (hist[a] != hist[b]) return hist[a] < hist[b];
        return a > b;
    });
    std::vector<uint64_t> w(syms.size());
    for (size_t i = 0; i < syms.size(); ++i) w[i] = hist[syms[i]];
    std::vector<uint32_t> len;
    if (!package_merge(w, max_len, len)) return HD_ERR_INVALID;
    // most frequent (shortest) first, like encoder/src/huff.cpp:189-194
    uint32_t k = 0;
    for (size_t i = syms.size(); i-- > 0;) {
        pairs_out[k].symbol = (uint8_t)syms[i];
        pairs_out[k].length = (uint8_t)len[i];
        ++k;
    }
    std::stable_sort(pairs_out, pairs_out + k, [](const hd_sym_len &a, const hd_sym_len &b) { return a.length < b.length; });
    *n_pairs_out = k;
    return HD_OK;
}

hd_status hd_flat_lut_from_lengths(const hd_sym_len *pairs, uint32_t n_pairs,
                                   uint32_t lut_bits, hd_lut_entry *lut_out) {
    HostCode c;
    if (!lut_out || lut_bits == 0 || lut_bits > kMaxCodeLen) return HD_ERR_INVALID;
    if (!code_from_pairs(pairs, n_pairs, c)) return HD_ERR_FORMAT;
    if (c.max_len > lut_bits) return HD_ERR_INVALID;
    std::memset(lut_out, 0, sizeof(hd_lut_entry) << lut_bits);
    for (int s = 0; s < 256; ++s) {
        uint32_t len = c.len[s];
        if (!len) continue;
        uint32_t span = 1u << (lut_bits - len);
        uint32_t base = c.code[s] << (lut_bits - len);
        for (uint32_t j = 0; j < span; ++j) { lut_out[base + j].num_bits = (uint8_t)len; lut_out[base + j].symbol = (uint8_t)s; }
    }
    return HD_OK;
}

static unsigned pick_threads(int n_threads, uint64_t n) {
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    if (nt > 256) nt = 256;
    uint64_t maxt = n / (1u << 16) + 1;   // at least 64 KiB per thread
    if (nt > maxt) nt = (unsigned)maxt;
    return nt;
}

hd_status hd_encoded_size(const uint8_t *data, uint64_t n, const hd_sym_len *pairs,
                          uint32_t n_pairs, uint64_t *bits_out) {
    HostCode c;
    if (!bits_out || (!data && n)) return HD_ERR_INVALID;
    if (!code_from_pairs(pairs, n_pairs, c)) return HD_ERR_FORMAT;
    unsigned nt = pick_threads(0, n);
    std::vector<uint64_t> part(nt, 0);
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            uint64_t b = n * t / nt, e = n * (t + 1) / nt, s = 0;
            for (uint64_t i = b; i < e; ++i) { uint32_t l = c.len[data[i]]; if (!l) bad = 1; s += l; }
            part[t] = s;
        });
    }
    for (auto &x : th) x.join();
    if (bad) return HD_ERR_INVALID;   // symbol without a code
    uint64_t bits = 0;
    for (auto v : part) bits += v;
    *bits_out = bits;
    return HD_OK;
}

// MSB-first packer: first code bit = bit 31 of unit 0
// (gpuhd/encoder/src/llhuffman_encoder.cc:216-231, encoder.cu:295-331); gap rule
// of encoder.cu:307-312: a codeword starting at bit p and ending at e with
// e/128 != p/128 records gap[p/128] = e & 15.
hd_status hd_encode(const uint8_t *data, uint64_t n, const hd_sym_len *pairs,
                    uint32_t n_pairs, uint32_t *units_out, uint64_t cap_units,
                    int pad_unit, uint32_t *gap_words_out, int n_threads) {
    HostCode c;
    if (!units_out || (!data && n)) return HD_ERR_INVALID;
    if (!code_from_pairs(pairs, n_pairs, c)) return HD_ERR_FORMAT;
    unsigned nt = pick_threads(n_threads, n);
    std::vector<uint64_t> start(nt + 1, 0);
    {
        std::vector<std::thread> th;
        std::atomic<int> bad{0};
        for (unsigned t = 0; t < nt; ++t) {
            th.emplace_back([&, t]() {
                uint64_t b = n * t / nt, e = n * (t + 1) / nt, s = 0;
                for (uint64_t i = b; i < e; ++i) { uint32_t l = c.len[data[i]]; if (!l) bad = 1; s += l; }
                start[t + 1] = s;
            });
        }
        for (auto &x : th) x.join();
        if (bad) return HD_ERR_INVALID;
    }
    for (unsigned t = 0; t < nt; ++t) start[t + 1] += start[t];
    const uint64_t bits = start[nt];
    const uint64_t n_units = (bits + 31) / 32;
    const uint64_t n_seg = (bits + 127) / 128;
    if (n_units + (pad_unit ? 1 : 0) > cap_units) return HD_ERR_INVALID;
    std::vector<uint8_t> gap_bytes;
    if (gap_words_out) gap_bytes.assign(n_seg + 1, 0);

    // every thread owns the units strictly inside its bit range; the (at most
    // two) units it shares with neighbours are merged after the join
    struct Edge { uint64_t unit; uint32_t value; };
    std::vector<Edge> edges(2 * nt, Edge{~0ull, 0});
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            uint64_t b = n * t / nt, e = n * (t + 1) / nt;
            uint64_t p = start[t];
            const uint64_t p_end = start[t + 1];
            if (p == p_end) return;
            const uint64_t first_unit = p >> 5;
            const uint64_t last_unit = (p_end - 1) >> 5;
            uint64_t acc = 0;            // bits being assembled, left-aligned in 64
            uint32_t fill = (uint32_t)(p & 31);   // bits already present in current unit
            uint64_t unit = first_unit;
            auto emit = [&](uint32_t v) {
                if (unit == first_unit && (start[t] & 31)) edges[2 * t] = Edge{unit, v};
                else if (unit == last_unit && (p_end & 31) && t + 1 < nt) edges[2 * t + 1] = Edge{unit, v};
                else units_out[unit] = v;
                ++unit;
            };
            for (uint64_t i = b; i < e; ++i) {
                const uint32_t len = c.len[data[i]];
                const uint64_t code = c.code[data[i]];
                if (gap_words_out) {
                    uint64_t en = p + len;
                    if ((en >> 7) != (p >> 7)) gap_bytes[p >> 7] = (uint8_t)(en & 15);
                    p = en;
                }
                acc |= code << (64 - fill - len);
                fill += len;
                if (fill >= 32) {
                    emit((uint32_t)(acc >> 32));
                    acc <<= 32;
                    fill -= 32;
                }
            }
            if (fill) emit((uint32_t)(acc >> 32));
        });
    }
    for (auto &x : th) x.join();
    // merge shared units (zero-initialise them first)
    for (auto &ed : edges) if (ed.unit != ~0ull) units_out[ed.unit] = 0;
    for (auto &ed : edges) if (ed.unit != ~0ull) units_out[ed.unit] |= ed.value;
    if (pad_unit) units_out[n_units] = 0;
    if (gap_words_out) {
        const uint64_t n_words = (n_seg + 7) / 8;
        for (uint64_t w = 0; w < n_words; ++w) {
            uint32_t v = 0;
            for (uint32_t j = 0; j < 8 && w * 8 + j < n_seg; ++j) v |= (uint32_t)(gap_bytes[w * 8 + j] & 15) << (4 * j);
            gap_words_out[w] = v;
        }
    }
    return HD_OK;
}

// ------------------------------------------------------------ .huff container
static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static void wr32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

size_t hd_huff_image_bytes(uint64_t symbol_count, uint32_t n_units, uint32_t n_segments) {
    return 8 + 2 * (size_t)symbol_count + 12 + 4 * (((size_t)n_segments + 7) / 8 + (size_t)n_units);
}

// layout: Huffman_coding_Gap_arrays/encoder/src/huff.cpp:186-202,
// reader decoder/src/huff.cpp:36-100
hd_status hd_huff_parse(const void *file, size_t size, hd_huff_view *v) {
    if (!file || !v) return HD_ERR_INVALID;
    const uint8_t *p = static_cast<const uint8_t *>(file);
    if (size < 8) return HD_ERR_FORMAT;
    uint64_t sc = (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32);
    if (sc == 0 || sc > 256) return HD_ERR_FORMAT;
    size_t off = 8;
    if (size < off + 2 * sc + 12) return HD_ERR_FORMAT;
    v->symbol_count = sc;
    v->symbols = reinterpret_cast<const hd_sym_len *>(p + off);
    off += 2 * sc;
    v->original_size = rd32(p + off);
    v->n_units = rd32(p + off + 4);
    v->n_segments = rd32(p + off + 8);
    off += 12;
    // the reference reader stores these in `int` (decoder/src/huff.cpp:81-86)
    if (v->original_size > 0x7fffffffu || v->n_units > 0x7fffffffu) return HD_ERR_FORMAT;
    if ((uint64_t)v->n_segments != ((uint64_t)v->n_units + 3) / 4) return HD_ERR_FORMAT;
    size_t gw = ((size_t)v->n_segments + 7) / 8;
    if (size < off + 4 * (gw + (size_t)v->n_units)) return HD_ERR_FORMAT;
    v->gap_words = p + off;
    v->units = p + off + 4 * gw;
    v->total_bytes = off + 4 * (gw + (size_t)v->n_units);
    uint32_t prev = 0;
    for (uint64_t i = 0; i < sc; ++i) {
        uint32_t l = v->symbols[i].length;
        if (l == 0 || l > kMaxCodeLen || l < prev) return HD_ERR_FORMAT;
        prev = l;
    }
    return HD_OK;
}

hd_status hd_huff_write(void *dst, size_t cap, const hd_sym_len *pairs, uint32_t n_pairs,
                        uint32_t original_size, const uint32_t *units, uint32_t n_units,
                        const uint32_t *gap_words, uint32_t n_segments, size_t *written) {
    if (!dst || !pairs || !n_pairs || n_pairs > 256 || (!units && n_units) || (!gap_words && n_segments))
        return HD_ERR_INVALID;
    size_t need = hd_huff_image_bytes(n_pairs, n_units, n_segments);
    if (cap < need) return HD_ERR_INVALID;
    uint8_t *p = static_cast<uint8_t *>(dst);
    wr32(p, n_pairs);
    wr32(p + 4, 0);
    size_t off = 8;
    for (uint32_t i = 0; i < n_pairs; ++i) { p[off++] = pairs[i].symbol; p[off++] = pairs[i].length; }
    wr32(p + off, original_size);
    wr32(p + off + 4, n_units);
    wr32(p + off + 8, n_segments);
    off += 12;
    size_t gw = ((size_t)n_segments + 7) / 8;
    for (size_t i = 0; i < gw; ++i, off += 4) wr32(p + off, gap_words[i]);
    for (size_t i = 0; i < n_units; ++i, off += 4) wr32(p + off, units[i]);
    if (written) *written = off;
    return HD_OK;
}

// ------------------------------------------------------------- synthetic data
// Distribution of generate.cpp:33-47.  The reference draws everything from one
// mt19937_64 seeded from random_device (generate.cpp:32, not reproducible); here
// every 1 MiB block has its own engine seeded from (seed, block) so blocks can
// be produced in parallel.
hd_status hd_generate(uint8_t *out, uint64_t n, double redundancy, uint64_t seed, int n_threads) {
    if (!out && n) return HD_ERR_INVALID;
    redundancy = std::max(0.0, std::min(1.0, redundancy));
    const uint64_t kBlock = 1u << 20;
    const uint64_t n_blocks = (n + kBlock - 1) / kBlock;
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    if (nt > n_blocks) nt = (unsigned)std::max<uint64_t>(1, n_blocks);
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        th.emplace_back([&]() {
            for (;;) {
                uint64_t b = next.fetch_add(1);
                if (b >= n_blocks) break;
                std::seed_seq sq{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)b, (uint32_t)(b >> 32), 0x48443230u};
                std::mt19937_64 rng(sq);
                std::uniform_int_distribution<int> dist_full(0, 255);
                std::uniform_int_distribution<int> dist_small(0, 3);
                uint64_t lo = b * kBlock, hi = std::min(n, lo + kBlock);
                for (uint64_t i = lo; i < hi; ++i) {
                    double r = std::generate_canonical<double, 10>(rng);
                    out[i] = (r < redundancy) ? (uint8_t)('A' + dist_small(rng)) : (uint8_t)dist_full(rng);
                }
            }
        });
    }
    for (auto &x : th) x.join();
    return HD_OK;
}

// -------------------------------------------------------------------- shards
void hd_shard_segments(uint64_t n_segments, uint32_t rank, uint32_t world,
                       uint64_t *seg_begin, uint64_t *seg_end) {
    if (world == 0) world = 1;
    const uint64_t tile = kTileThreads;
    const uint64_t n_tiles = (n_segments + tile - 1) / tile;
    uint64_t b = n_tiles * rank / world * tile;
    uint64_t e = n_tiles * (rank + 1ull) / world * tile;
    if (b > n_segments) b = n_segments;
    if (e > n_segments || rank + 1 == world) e = n_segments;
    if (seg_begin) *seg_begin = b;
    if (seg_end) *seg_end = e;
}

}  // extern "C"
