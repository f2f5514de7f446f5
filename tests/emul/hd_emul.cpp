// hd_emul.cpp -- TEST INFRASTRUCTURE.  Single-steps, on the CPU, the very same
// per-thread source the CUDA kernels run (csrc/hd_core.h) with the kernels'
// block/tile structure re-enacted sequentially, so that control-flow and table
// bugs are caught in this GPU-less container.  Never linked into the product
// library; the product has no CPU path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "hd_core.h"
#include "hd_internal.h"

using namespace hd;

namespace {

struct HostTables {
    TableBlob blob;
    DecodeTables T;
    SyncLut L;
};

void bind(HostTables &h) {
    const uint8_t *b = h.blob.bytes.data();
    h.T.lut = reinterpret_cast<const uint32_t *>(b);
    h.T.idx_shift = 32 - h.blob.lut_bits;
    h.T.lut_bits = h.blob.lut_bits;
    h.T.len_tab = b + h.blob.off_len;
    h.T.order = b + h.blob.off_order;
    h.T.lt = reinterpret_cast<const LongTables *>(b + h.blob.off_long);
    h.L.lut = reinterpret_cast<const uint16_t *>(b + h.blob.off_sync);
    h.L.idx_shift = 32 - h.blob.lut_bits;
    h.L.lm = h.blob.lut_bits;
    h.L.lt = h.T.lt;
    h.L.order = h.T.order;
    h.L.has_long = h.blob.has_long != 0;
}

bool make_tables(const hd_lut_entry *flat, uint32_t flat_bits, const hd_sym_len *pairs, uint32_t n_pairs,
                 HostTables &h) {
    HostCode c;
    bool ok = flat ? code_from_flat_lut(flat, flat_bits, c) : code_from_pairs(pairs, n_pairs, c);
    if (!ok) return false;
    build_table_blob(c, h.blob);
    bind(h);
    return true;
}

struct HostSegLoader {
    const uint32_t *units;
    uint64_t n_units;
    Seg4 operator()(uint64_t s) const {
        Seg4 r{0, 0, 0, 0};
        const uint64_t u = s * 4;
        if (u < n_units) r.x = units[u];
        if (u + 1 < n_units) r.y = units[u + 1];
        if (u + 2 < n_units) r.z = units[u + 2];
        if (u + 3 < n_units) r.w = units[u + 3];
        return r;
    }
};

}  // namespace

// Re-enactment of hd_decode_write_kernel<NT, SEGS>: tiles of NT chunks of `segs`
// segments; per thread: transposed input column, decode_chunk -> staging column
// (same address arithmetic as on the device, over a host arena whose staging
// region starts on a row boundary), exclusive scan, clamp, compact_to_tile into
// a tile placed at the same 16-byte phase, copy out.
template <int NT>
static int64_t emul_decode_gap_nt(HostTables &h, const uint32_t *units, uint64_t n_units, const uint32_t *gap_words,
                                  uint64_t n_segments, uint32_t first_bit, uint8_t *out, uint64_t n_symbols,
                                  uint64_t out_addr_phase, uint32_t segs) {
    const uint32_t min_len = h.blob.min_len ? h.blob.min_len : 1;
    const uint32_t max_cnt = (kSegmentBits * segs + 15 + min_len - 1) / min_len;
    const uint32_t stage_words = max_cnt / 4 + 3;
    const uint32_t tile_sym = (NT * kSegmentBits * segs + 15 + min_len - 1) / min_len + 2;
    const uint32_t tile_bytes = ((tile_sym + 32) + 15) & ~15u;
    const uint32_t in_rows = 4 * segs + 3;
    const uint32_t row = NT * 4;
    std::vector<uint32_t> in((size_t)in_rows * NT);
    std::vector<uint8_t> arena((size_t)stage_words * row + 2 * row);
    const uint32_t stage_off = row;   // row-aligned offset inside the arena (device: aligned shared address)
    std::vector<uint8_t> tile(tile_bytes);
    auto nib = [&](uint64_t j) { return (gap_words[j >> 3] >> (4u * (uint32_t)(j & 7))) & 15u; };
    const uint64_t n_chunks = (n_segments + segs - 1) / segs;
    const uint64_t n_tiles = (n_chunks + NT - 1) / NT;
    std::vector<uint32_t> cnt(NT);
    uint64_t base = 0;
    for (uint64_t tl = 0; tl < n_tiles; ++tl) {
        // poison so that reads of never-written staging/tile bytes show up
        std::fill(arena.begin(), arena.end(), (uint8_t)0xDE);
        std::fill(tile.begin(), tile.end(), (uint8_t)0xEE);
        std::fill(in.begin(), in.end(), 0xBADC0DEu);
        for (uint32_t t = 0; t < NT; ++t) {
            const uint64_t g0 = (tl * NT + t) * segs;
            cnt[t] = 0;
            if (g0 >= n_segments) continue;
            const uint64_t left = n_segments - g0;
            const uint32_t nseg = left < segs ? (uint32_t)left : segs;
            uint32_t *col = in.data() + t;
            for (uint32_t w = 0; w < 4 * segs + 2; ++w) {
                const uint64_t u = g0 * 4 + w;
                col[(size_t)w * NT] = u < n_units ? units[u] : 0u;
            }
            const uint32_t entry = g0 == 0 ? first_bit : nib(g0 - 1);
            const uint32_t stop = nseg * kSegmentBits + nib(g0 + nseg - 1);
            const uint32_t sa = stage_off + t * 4u;
            uint32_t c = h.blob.has_long ? decode_chunk<NT, true>(col, entry, stop, h.T, sa, arena.data())
                                         : decode_chunk<NT, false>(col, entry, stop, h.T, sa, arena.data());
            if (c > max_cnt) return -2;   // staging/tile sizing assumption violated
            if (c / 4 + 1 >= stage_words) return -5;   // compact_to_tile would read past the staging rows
            cnt[t] = c;
        }
        // canary: the row after the staging rows must be untouched (a byte store past the last row lands there)
        for (uint32_t i = 0; i < row; ++i)
            if (arena[(size_t)stage_off + (size_t)stage_words * row + i] != 0xDE) return -6;
        const uint32_t *stage = reinterpret_cast<const uint32_t *>(arena.data() + stage_off);
        uint32_t excl = 0;
        const uint32_t A = (uint32_t)((out_addr_phase + base) & 15u);
        uint64_t tile_total = 0;
        for (uint32_t t = 0; t < NT; ++t) {
            uint64_t my_off = base + excl;
            uint32_t c = cnt[t];
            if (my_off >= n_symbols) c = 0;
            else if (my_off + c > n_symbols) c = (uint32_t)(n_symbols - my_off);
            if (c) {
                if (A + excl + c > tile_bytes) return -3;
                compact_to_tile(stage + t, NT, c, tile.data(), A + excl);
            }
            excl += cnt[t];
            tile_total += cnt[t];
        }
        uint64_t lim = base >= n_symbols ? 0 : n_symbols - base;
        uint64_t tot = tile_total < lim ? tile_total : lim;
        std::memcpy(out + base, tile.data() + A, tot);
        base += tile_total;
    }
    return (int64_t)(base < n_symbols ? base : n_symbols);
}

extern "C" {

int64_t emul_decode_gap(const uint32_t *units, uint64_t n_units, const uint32_t *gap_words, uint64_t n_segments,
                        uint32_t first_bit, const hd_lut_entry *flat, uint32_t flat_bits, const hd_sym_len *pairs,
                        uint32_t n_pairs, uint8_t *out, uint64_t n_symbols, uint64_t out_addr_phase, uint32_t segs,
                        uint32_t nt) {
    HostTables h;
    if (!make_tables(flat, flat_bits, pairs, n_pairs, h)) return -1;
    if (segs != 1 && segs != 2 && segs != 4) return -4;
    if (nt == 256) return emul_decode_gap_nt<256>(h, units, n_units, gap_words, n_segments, first_bit, out, n_symbols, out_addr_phase, segs);
    if (nt == 512) return emul_decode_gap_nt<512>(h, units, n_units, gap_words, n_segments, first_bit, out, n_symbols, out_addr_phase, segs);
    return -4;
}

// Re-enactment of hd_sync_chunks_kernel + hd_sync_seams_kernel rounds.
// Returns the number of seam rounds run (>= 0) or negative on error.
// rounds_out[0] = K1a intra-block iterations (max over blocks).
int64_t emul_sync(const uint32_t *units, uint64_t n_units, uint32_t first_bit, const hd_lut_entry *flat,
                  uint32_t flat_bits, const hd_sym_len *pairs, uint32_t n_pairs, uint32_t *gap_words_out,
                  uint32_t *rounds_out) {
    HostTables h;
    if (!make_tables(flat, flat_bits, pairs, n_pairs, h)) return -1;
    const uint64_t n_seg = (n_units + 3) / 4;
    const uint64_t n_chunks = (n_seg + kChunkSegs - 1) / kChunkSegs;
    const uint64_t n_blocks = (n_chunks + kSyncThreads - 1) / kSyncThreads;
    HostSegLoader ld{units, n_units};
    std::vector<uint8_t> blk_entry(n_blocks, 0);
    uint32_t max_iters = 0;
    // ---- K1a ----
    for (uint64_t b = 0; b < n_blocks; ++b) {
        uint32_t word[kSyncThreads] = {}, entry[kSyncThreads] = {}, s_exit[kSyncThreads] = {};
        bool active[kSyncThreads];
        bool merged;
        for (uint32_t tid = 0; tid < kSyncThreads; ++tid) {
            const uint64_t chunk = b * kSyncThreads + tid;
            active[tid] = chunk < n_chunks;
            entry[tid] = chunk == 0 ? first_bit : 0u;
            if (active[tid]) word[tid] = sync_parse_chunk(ld, chunk, entry[tid], h.L, false, 0u, merged);
            s_exit[tid] = word[tid] >> 28;
        }
        uint32_t iters = 0;
        for (;;) {
            ++iters;
            uint32_t xin[kSyncThreads];
            bool any = false;
            for (uint32_t tid = 0; tid < kSyncThreads; ++tid)
                xin[tid] = (tid == 0 || !active[tid]) ? entry[tid] : s_exit[tid - 1];
            for (uint32_t tid = 0; tid < kSyncThreads; ++tid) {
                if (xin[tid] == entry[tid]) continue;
                any = true;
                const uint64_t chunk = b * kSyncThreads + tid;
                word[tid] = sync_parse_chunk(ld, chunk, xin[tid], h.L, true, word[tid], merged);
                entry[tid] = xin[tid];
                s_exit[tid] = word[tid] >> 28;
            }
            if (!any) break;
        }
        if (iters > max_iters) max_iters = iters;
        for (uint32_t tid = 0; tid < kSyncThreads; ++tid)
            if (active[tid]) gap_words_out[b * kSyncThreads + tid] = word[tid];
        blk_entry[b] = (uint8_t)entry[0];
    }
    if (rounds_out) rounds_out[0] = max_iters;
    // ---- K1b rounds until no block exit changes ----
    int64_t rounds = 0;
    for (;; ++rounds) {
        uint32_t changed = 0;
        for (uint64_t b = 1; b < n_blocks; ++b) {
            const uint64_t c0 = b * kSyncThreads;
            const uint32_t xin = gap_words_out[c0 - 1] >> 28;
            if (xin == blk_entry[b]) continue;
            uint64_t c_end = c0 + kSyncThreads;
            if (c_end > n_chunks) c_end = n_chunks;
            uint32_t entry = xin;
            bool merged = false;
            for (uint64_t c = c0; c < c_end; ++c) {
                uint32_t w = sync_parse_chunk(ld, c, entry, h.L, true, gap_words_out[c], merged);
                gap_words_out[c] = w;
                if (merged) break;
                entry = w >> 28;
            }
            blk_entry[b] = (uint8_t)xin;
            if (!merged) ++changed;
        }
        if (!changed) break;
        if ((uint64_t)rounds > n_blocks + 2) return -2;
    }
    return rounds + 1;
}

// table introspection for tests
int emul_table_info(const hd_lut_entry *flat, uint32_t flat_bits, const hd_sym_len *pairs, uint32_t n_pairs,
                    uint32_t info[4]) {
    HostTables h;
    if (!make_tables(flat, flat_bits, pairs, n_pairs, h)) return -1;
    info[0] = h.blob.lut_bits;
    info[1] = h.blob.min_len;
    info[2] = h.blob.max_len;
    info[3] = h.blob.has_long;
    return 0;
}

}  // extern "C"
