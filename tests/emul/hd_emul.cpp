// hd_emul.cpp -- TEST INFRASTRUCTURE.  Single-steps, on the CPU, the very same
// per-thread source the CUDA kernels run (csrc/hd_core.h) with the kernels'
// block/tile structure re-enacted sequentially, so that control-flow and table
// bugs are caught in this GPU-less container.  Never linked into the product
// library; the product has no CPU path.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "hd_core.h"
#include "hd_internal.h"

using namespace hd;

namespace {

constexpr uint32_t kSyncWalkChunksEmul = hd::kSyncTileChunks;   // == kSyncWalkChunks of hd_kernels.cuh (device header)

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
    h.T.one = 1;
    h.T.sel_len = hd::kSelLen;
    h.T.sel_cnt8 = hd::kSelCnt8;
    h.T.pair_slope = hd::kPairSelSlope;
    h.T.len_tab = b + h.blob.off_len;
    h.T.order = b + h.blob.off_order;
    h.T.lt = reinterpret_cast<const LongTables *>(b + h.blob.off_long);
    h.L.lut = reinterpret_cast<const uint16_t *>(b + h.blob.off_sync);
    h.L.idx_shift = 32 - h.blob.lut_bits;
    h.L.lm = h.blob.lut_bits;
    h.L.lt = h.T.lt;
    h.L.order = h.T.order;
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

// Re-enactment of hd_decode_tiles_kernel<SEGS, RCAP> (hd_decode.cuh): warp tiles of 32 chunks; per lane: transposed
// input column, decode_chunk<32> into the warp's staging block, then the IN-PLACE compaction exactly as the kernel
// orders it (all lanes column_read -> all lanes column_write -> all lanes column_write_head) at the 16-byte phase of
// the warp's destination, copy out.
template <int RCAP>
static int64_t emul_decode_tiles(HostTables &h, const uint32_t *units, uint64_t n_units, const uint32_t *gap_words,
                                 uint64_t n_segments, uint32_t first_bit, uint8_t *out, uint64_t n_symbols,
                                 uint64_t out_addr_phase, uint32_t segs) {
    const uint32_t min_len = h.blob.min_len ? h.blob.min_len : 1;
    const uint32_t max_cnt = (kSegmentBits * segs + 15 + min_len - 1) / min_len;
    const uint32_t stage_rows = max_cnt / 4 + 3;
    if ((max_cnt + 3) / 4 > (uint32_t)RCAP) return -4;
    const uint32_t in_rows = 4 * segs + 2;
    std::vector<uint32_t> in((size_t)in_rows * 32);
    std::vector<uint8_t> arena((size_t)(stage_rows + 2) * 128);
    const uint32_t stage_off = 128;
    auto nib = [&](uint64_t j) { return (gap_words[j >> 3] >> (4u * (uint32_t)(j & 7))) & 15u; };
    const uint64_t n_chunks = (n_segments + segs - 1) / segs;
    const uint64_t n_wtiles = (n_chunks + 31) / 32;
    uint64_t base = 0;
    for (uint64_t wt = 0; wt < n_wtiles; ++wt) {
        std::fill(arena.begin(), arena.end(), (uint8_t)0xDE);
        std::fill(in.begin(), in.end(), 0xBADC0DEu);
        uint32_t cnt[32];
        for (uint32_t l = 0; l < 32; ++l) {
            const uint64_t g0 = (wt * 32 + l) * segs;
            cnt[l] = 0;
            if (g0 >= n_segments) continue;
            const uint64_t left = n_segments - g0;
            const uint32_t nseg = left < segs ? (uint32_t)left : segs;
            uint32_t *col = in.data() + l;
            for (uint32_t w = 0; w < in_rows; ++w) {
                const uint64_t u = g0 * 4 + w;
                col[(size_t)w * 32] = u < n_units ? units[u] : 0u;
            }
            const uint32_t entry = g0 == 0 ? first_bit : nib(g0 - 1);
            const uint32_t stop = nseg * kSegmentBits + nib(g0 + nseg - 1);
            const uint32_t sa = stage_off + l * 4u;
            const uint32_t c = h.blob.has_long ? decode_chunk<32, true>(col, entry, stop, h.T, sa, arena.data())
                                               : decode_chunk<32, false>(col, entry, stop, h.T, sa, arena.data());
            if (c > max_cnt) return -2;
            cnt[l] = c;
        }
        for (uint32_t i = 0; i < 128; ++i)   // canary row behind the staging rows
            if (arena[(size_t)stage_off + (size_t)stage_rows * 128 + i] != 0xDE) return -6;
        uint8_t *sblk = arena.data() + stage_off;
        uint32_t wtot = 0, rows_max = 0;
        for (uint32_t l = 0; l < 32; ++l) { wtot += cnt[l]; rows_max = std::max(rows_max, (cnt[l] + 3) / 4); }
        const uint64_t lim = base >= n_symbols ? 0 : n_symbols - base;
        const uint32_t total = wtot < lim ? wtot : (uint32_t)lim;
        const uint32_t phase = (uint32_t)((out_addr_phase + base) & 15u);
        if (phase + wtot + 4 > stage_rows * 128u) return -3;   // compacted span + surplus of the last word must fit
        uint32_t w[32][RCAP + 2], c[32], excl[32];
        uint32_t e = 0;
        for (uint32_t l = 0; l < 32; ++l) {
            excl[l] = e;
            e += cnt[l];
            c[l] = cnt[l];
            if (excl[l] >= total) c[l] = 0;
            else if (excl[l] + c[l] > total) c[l] = total - excl[l];
            for (int i = 0; i < RCAP + 2; ++i) w[l][i] = 0;
            column_read<RCAP>(stage_off + l * 4u, rows_max, w[l], arena.data());
        }
        for (uint32_t l = 0; l < 32; ++l) column_write<RCAP>(w[l], c[l], stage_off + phase + excl[l], rows_max, arena.data());
        for (uint32_t l = 0; l < 32; ++l) column_write_head(w[l][0], c[l], stage_off + phase + excl[l], arena.data());
        std::memcpy(out + base, sblk + phase, total);
        base += wtot;
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
    if (nt == 32) {
        int64_t r = emul_decode_tiles<12>(h, units, n_units, gap_words, n_segments, first_bit, out, n_symbols, out_addr_phase, segs);
        if (r == -4) r = emul_decode_tiles<24>(h, units, n_units, gap_words, n_segments, first_bit, out, n_symbols, out_addr_phase, segs);
        if (r == -4 && segs == 1) r = emul_decode_tiles<36>(h, units, n_units, gap_words, n_segments, first_bit, out, n_symbols, out_addr_phase, segs);
        return r == -4 ? -40 : r;   // -40: geometry not instantiated for this code (mirrors tile_geom in hd_api.cu)
    }
    return -4;
}

}  // extern "C" (closed for the templates below)

// Re-enactment of hd_sync_parse_kernel (K1a: halo + chunk per thread, seams
// settled inside the warp tile, tile seams checked against the published exits)
// and, if a published exit moved afterwards, hd_sync_walk_kernel rounds (K1b).
// Returns the number of K1b rounds run (>= 0) or negative on error.
// rounds_out[0] = tiles whose published exit moved, [1] = K1b rounds.
template <bool EXACT, bool LONG>
static int64_t emul_sync_t(HostTables &h, const uint32_t *units, uint64_t n_units, uint32_t first_bit,
                           uint32_t *gap_words_out, uint32_t *rounds_out) {
    const uint64_t n_seg = (n_units + 3) / 4;
    const uint64_t n_chunks = (n_seg + kSyncChunkSegs - 1) / kSyncChunkSegs;
    const uint32_t kW = kSyncChunkSegs / 8;
    HostSegLoader ld{units, n_units};
    std::vector<uint8_t> cur_entry(n_chunks, 0);
    // ---- K1a: speculative halo+chunk parse, then the in-warp-tile seam rounds ----
    std::vector<uint32_t> ent(n_chunks, 0);
    for (uint64_t c = 0; c < n_chunks; ++c) {
        uint64_t seg = c * kSyncChunkSegs;
        uint32_t pos = 0;
        if (c == 0) {
            pos = first_bit;
        } else {
            for (uint64_t sg = seg - kSyncHaloSegs; sg < seg; ++sg) {
                Seg4 w = ld(sg), n = ld(sg + 1);
                pos = sync_segment<EXACT, LONG>(w.x, w.y, w.z, w.w, n.x, pos, h.L);
            }
        }
        ent[c] = pos;
        for (uint32_t g = 0; g < kW; ++g) {
            uint32_t word = 0;
            for (uint32_t k = 0; k < 8; ++k, ++seg) {
                Seg4 w = ld(seg), n = ld(seg + 1);
                pos = sync_segment<EXACT, LONG>(w.x, w.y, w.z, w.w, n.x, pos, h.L);
                if (pos > 15) return -3;
                word |= pos << (4u * k);
            }
            gap_words_out[c * kW + g] = word;
        }
    }
    // one warp tile = chunks [31t, 31t+32); lane 0 (chunk 31t, t > 0) is owned by the previous tile: its parse runs
    // here on private copies (from its own halo) only to give lane 1 an exit offset
    std::vector<uint32_t> fin_words(gap_words_out, gap_words_out + n_chunks * kW);   // spec results of every chunk
    std::vector<uint32_t> fin_ent(ent);
    const uint64_t n_tiles_k1 = n_chunks > 1 ? (n_chunks - 1 + kSyncTileChunks - 1) / kSyncTileChunks : 1;
    std::vector<uint32_t> published(n_tiles_k1, 0), assumed(n_tiles_k1, 0);
    for (uint64_t t = 0; t < n_tiles_k1; ++t) {
        const uint64_t t0 = t * kSyncTileChunks;
        const uint64_t t1 = t0 + 32 < n_chunks ? t0 + 32 : n_chunks;
        uint32_t w[32][2], e[32];
        for (uint64_t c = t0; c < t1; ++c) {   // every lane starts from its speculative (halo) parse
            w[c - t0][0] = gap_words_out[c * kW];
            w[c - t0][1] = gap_words_out[c * kW + 1];
            e[c - t0] = ent[c];
        }
        for (;;) {
            uint32_t left[32];
            bool any = false;
            for (uint64_t l = 1; l < t1 - t0; ++l) left[l] = w[l - 1][1] >> 28;   // shuffle
            for (uint64_t l = 1; l < t1 - t0; ++l) {
                if (left[l] == e[l]) continue;
                any = true;
                sync_repair_chunk<EXACT, LONG>(ld, t0 + l, left[l], h.L, w[l]);
                e[l] = left[l];
            }
            if (!any) break;
        }
        assumed[t] = w[0][1] >> 28;                        // lane 0's exit: what lane 1 was entered with
        if (t1 - t0 == 32) published[t] = w[31][1] >> 28;   // exit of the tile's last chunk
        for (uint64_t l = (t == 0 ? 0 : 1); l < t1 - t0; ++l) {   // owners write
            fin_words[(t0 + l) * kW] = w[l][0];
            fin_words[(t0 + l) * kW + 1] = w[l][1];
            fin_ent[t0 + l] = e[l];
        }
    }
    for (uint64_t i = 0; i < n_chunks * kW; ++i) gap_words_out[i] = fin_words[i];
    ent = fin_ent;
    for (uint64_t c = 0; c < n_chunks; ++c) cur_entry[c] = (uint8_t)ent[c];
    // ---- the seams between warp tiles, inside K1a: tile t assumed lane 0's speculative exit of chunk 31t, tile t-1
    // PUBLISHED the real one right after settling its own seams (before any later repair of its own first seam: the
    // earliest value a concurrent warp can read).  A mismatch is repaired by a walk through the tile; a walk that does
    // not merge inside a full tile moves an exit that is already published and counts as changed.
    uint32_t changed0 = 0;
    for (uint64_t t = 1; t < n_tiles_k1; ++t) {
        const uint64_t c0 = t * kSyncTileChunks + 1;
        uint64_t c1 = c0 + kSyncTileChunks;
        if (c1 > n_chunks) c1 = n_chunks;
        uint32_t xin = published[t - 1];
        if (xin == assumed[t]) continue;
        bool exit_changed = false;
        for (uint64_t c = c0; c < c1; ++c) {
            uint32_t words[2] = {gap_words_out[c * kW], gap_words_out[c * kW + 1]};
            const bool merged = sync_repair_chunk<EXACT, LONG>(ld, c, xin, h.L, words);
            gap_words_out[c * kW] = words[0];
            gap_words_out[c * kW + 1] = words[1];
            cur_entry[c] = (uint8_t)xin;
            exit_changed = !merged;
            if (merged) break;
            xin = words[1] >> 28;
        }
        if (exit_changed && c1 - c0 == kSyncTileChunks) ++changed0;
    }
    // ---- K1b rounds, only if a published exit moved: until no group exit changes ----
    if (rounds_out) { rounds_out[0] = changed0; rounds_out[1] = 0; }
    if (!changed0) return 0;
    int64_t rounds = 0;
    const uint64_t n_groups = (n_chunks + kSyncWalkChunksEmul - 1) / kSyncWalkChunksEmul;
    for (uint64_t r = 0; r <= n_groups + 2; ++r) {
        uint32_t ch = 0;
        std::vector<uint8_t> exit_before(n_chunks);
        for (uint64_t c = 0; c < n_chunks; ++c) exit_before[c] = (uint8_t)(gap_words_out[c * kW + kW - 1] >> 28);
        for (uint64_t g = 0; g < n_groups; ++g) {
            const uint64_t c0 = g * kSyncWalkChunksEmul + 1;
            if (c0 >= n_chunks) continue;
            uint64_t c1 = c0 + kSyncWalkChunksEmul;
            if (c1 > n_chunks) c1 = n_chunks;
            uint32_t xin = exit_before[c0 - 1];
            bool exit_changed = false;
            for (uint64_t c = c0; c < c1; ++c) {
                uint32_t words[2] = {gap_words_out[c * kW], gap_words_out[c * kW + 1]};
                exit_changed = false;
                if (xin != cur_entry[c]) {
                    const bool merged = sync_repair_chunk<EXACT, LONG>(ld, c, xin, h.L, words);
                    gap_words_out[c * kW] = words[0];
                    gap_words_out[c * kW + 1] = words[1];
                    cur_entry[c] = (uint8_t)xin;
                    exit_changed = !merged;
                }
                if (!exit_changed) break;
                xin = words[1] >> 28;
            }
            if (exit_changed) ++ch;
        }
        ++rounds;
        if (rounds_out) rounds_out[1] += 1;
        if (!ch) return rounds;
    }
    return -2;
}

extern "C" int64_t emul_sync(const uint32_t *units, uint64_t n_units, uint32_t first_bit, const hd_lut_entry *flat,
                             uint32_t flat_bits, const hd_sym_len *pairs, uint32_t n_pairs, uint32_t *gap_words_out,
                             uint32_t *rounds_out, uint32_t exact) {
    HostTables h;
    if (!make_tables(flat, flat_bits, pairs, n_pairs, h)) return -1;
    const bool lng = h.blob.has_long != 0;
    if (exact) return lng ? emul_sync_t<true, true>(h, units, n_units, first_bit, gap_words_out, rounds_out)
                          : emul_sync_t<true, false>(h, units, n_units, first_bit, gap_words_out, rounds_out);
    return lng ? emul_sync_t<false, true>(h, units, n_units, first_bit, gap_words_out, rounds_out)
               : emul_sync_t<false, false>(h, units, n_units, first_bit, gap_words_out, rounds_out);
}

extern "C" {

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
