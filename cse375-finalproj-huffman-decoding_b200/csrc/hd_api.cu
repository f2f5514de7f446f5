// hd_api.cu -- C-ABI entry points (include/hd_b200.h) over the sm_100a kernels.
// Host orchestration only: workspace carving, launches, the self-sync fixed
// point loop.  Reference citations are relative to the reference repo root.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "hd_fixed.cuh"
#include "hd_kernels.cuh"
#include "hd_decode.cuh"
#include "hd_tables.cuh"
#include "hd_encoder.cuh"
#include "hd_mailbox.cuh"

using namespace hd;

namespace {

thread_local cudaError_t g_last_cuda = cudaSuccess;
std::atomic<uint64_t> g_launches{0};   // kernels launched by this library (hd_launch_count)

#define HD_CUDA(call)                              \
    do {                                           \
        cudaError_t e__ = (call);                  \
        if (e__ != cudaSuccess) {                  \
            g_last_cuda = e__;                     \
            (void)cudaGetLastError();              \
            return HD_ERR_CUDA;                    \
        }                                          \
    } while (0)

// ---- optional per-kernel device timing (hd_profile_*) -----------------------
// bench.py turns this on for a separate pass to obtain the dominant kernel's
// average launch duration on the launching stream; off by default (no events).
struct Profiler {
    bool on = false;
    cudaEvent_t ev[HD_PROFILE_SLOTS][2] = {};
    bool used[HD_PROFILE_SLOTS] = {};
    bool created = false;
};
thread_local Profiler g_prof;

struct ProfScope {
    int slot;
    cudaStream_t s;
    bool active;
    ProfScope(int slot_, cudaStream_t s_) : slot(slot_), s(s_), active(g_prof.on && !g_prof.used[slot_]) {
        if (active) cudaEventRecord(g_prof.ev[slot][0], s);
    }
    ~ProfScope() {
        if (active) {
            cudaEventRecord(g_prof.ev[slot][1], s);
            g_prof.used[slot] = true;
        }
    }
};

inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- workspace layout -------------------------------------------------------
// [ctrl A: 64 B {tile_counter u32}] [tile_status: n_tiles u64]   <- zeroed per K4 launch
// [ctrl B: 64 B {changed[8] u32, count u64}]                      <- zeroed per self-sync call
// [gap words (self-sync only): 2 x n_chunks u32]
// [cur_entry (self-sync only): n_chunks u8]
struct Workspace {
    uint32_t *tile_counter;
    uint64_t *tile_status;
    uint32_t *changed;       // 8 counters
    uint64_t *count;
    uint32_t *sync_exit;     // self-sync: one published exit word per K1a warp tile (zeroed per call)
    unsigned long long *tile_map;   // self-sync fallback: transfer function of every K1a tile (16 nibbles)
    uint8_t *tile_entry;            // self-sync fallback: exact entry offset of every K1a tile
    size_t sync_zero_bytes;  // self-sync: bytes from `changed` on that must be zeroed per call
    uint32_t *gap_words;
    uint8_t *blk_entry;      // per sync chunk: entry offset its landings were parsed from
    size_t zero_bytes;       // prefix of the workspace that must be zeroed per K4 launch
};

constexpr size_t kCtrlBytes = 64;

size_t ws_bytes(uint64_t n_units, bool selfsync) {
    const uint64_t n_seg = div_up(n_units, kSegmentUnits);
    const uint64_t n_tiles = std::max<uint64_t>(1, div_up(n_seg, kTileThreads));
    size_t b = kCtrlBytes + align_up(n_tiles * 8 * kStatusStride, 64) + kCtrlBytes;
    if (selfsync) {
        const uint64_t n_chunks = std::max<uint64_t>(1, div_up(n_seg, kSyncChunkSegs));
        b += align_up((div_up(n_chunks, kSyncTileChunks) + 1) * 4, 64);
        b += align_up(n_chunks * (kSyncChunkSegs / 8) * 4, 64) + align_up(n_chunks, 64);
        const uint64_t n_sync_tiles = div_up(n_chunks, kSyncTileChunks) + 1;
        b += align_up(n_sync_tiles * 8, 64) + align_up(n_sync_tiles, 64);
    }
    return b;
}

Workspace carve(void *ws, uint64_t n_units, bool selfsync) {
    Workspace w;
    uint8_t *p = static_cast<uint8_t *>(ws);
    const uint64_t n_seg = div_up(n_units, kSegmentUnits);
    const uint64_t n_tiles = std::max<uint64_t>(1, div_up(n_seg, kTileThreads));
    w.tile_counter = reinterpret_cast<uint32_t *>(p);
    w.tile_status = reinterpret_cast<uint64_t *>(p + kCtrlBytes);
    w.zero_bytes = kCtrlBytes + align_up(n_tiles * 8 * kStatusStride, 64);
    uint8_t *b = p + w.zero_bytes;
    w.changed = reinterpret_cast<uint32_t *>(b);
    w.count = reinterpret_cast<uint64_t *>(b + 48);
    w.gap_words = nullptr;
    w.blk_entry = nullptr;
    w.sync_exit = nullptr;
    w.tile_map = nullptr;
    w.tile_entry = nullptr;
    w.sync_zero_bytes = kCtrlBytes;
    if (selfsync) {
        const uint64_t n_chunks = std::max<uint64_t>(1, div_up(n_seg, kSyncChunkSegs));
        const size_t exit_bytes = align_up((div_up(n_chunks, kSyncTileChunks) + 1) * 4, 64);
        w.sync_exit = reinterpret_cast<uint32_t *>(b + kCtrlBytes);
        w.sync_zero_bytes = kCtrlBytes + exit_bytes;
        w.gap_words = reinterpret_cast<uint32_t *>(b + kCtrlBytes + exit_bytes);
        w.blk_entry = b + kCtrlBytes + exit_bytes + align_up(n_chunks * (kSyncChunkSegs / 8) * 4, 64);
        const uint64_t n_sync_tiles = div_up(n_chunks, kSyncTileChunks) + 1;
        w.tile_map = reinterpret_cast<unsigned long long *>(w.blk_entry + align_up(n_chunks, 64));
        w.tile_entry = reinterpret_cast<uint8_t *>(w.tile_map) + align_up(n_sync_tiles * 8, 64);
    }
    return w;
}

TableDev table_dev(const hd_table *t) {
    TableDev d;
    d.blob = t->d_blob;
    d.lut_bits = t->host.lut_bits;
    d.has_long = t->host.has_long;
    d.off_sync = t->host.off_sync;
    d.off_len = t->host.off_len;
    d.off_order = t->host.off_order;
    d.off_long = t->host.off_long;
    return d;
}

struct DeviceInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
};
DeviceInfo device_info() {
    static std::mutex mu;
    static DeviceInfo cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && cache[dev].sm_count) return cache[dev];
    DeviceInfo d;
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (dev >= 0 && dev < 64) cache[dev] = d;
    return d;
}

size_t sync_smem_bytes(const hd_table *t) {
    return align_up(2u << t->host.lut_bits, 16) + 256 + align_up(sizeof(LongTables), 16);
}

// ---- K4 "warp tile" kernel (hd_decode.cuh): geometry + launch ---------------------------------------------------
// segs = segments per decode thread, rcap = register rows of the in-place compaction (kernel template parameters),
// stage_rows = rows of one staging buffer of a warp block, n_warps = decode warps per CTA.
struct TileGeom {
    int segs = 0, rcap = 0;
    uint32_t stage_rows = 0, n_warps = 0;
    size_t smem = 0;
};
constexpr uint32_t kTileMaxWarps = 27;    // (27 + 1 scan warp) * 32 = 896 threads: 72 registers per thread
constexpr size_t kTileStaticSmem = 1024;  // static __shared__ of the kernel + alignment slack of the warp blocks

bool tile_geom(const hd_table *t, int segs, int max_smem, uint32_t want_warps, TileGeom &g) {
    const uint32_t min_len = std::max<uint32_t>(1, t->host.min_len);
    // a chunk decodes bits [entry, 128*segs + next_entry), offsets <= 15; decode_chunk may store one symbol more than
    // it counts (second symbol of the last lookup) and always writes the second byte of a lookup
    const uint32_t max_cnt = (kSegmentBits * segs + 15 + min_len - 1) / min_len;
    const uint32_t rows_used = (max_cnt + 3) / 4;
    g.segs = segs;
    g.stage_rows = max_cnt / 4 + 3;
    g.rcap = rows_used <= 12 ? 12 : rows_used <= 24 ? 24 : rows_used <= 36 ? 36 : 0;
    if (!g.rcap) return false;
    if (segs == 4 && g.rcap != 24) return false;   // not instantiated
    const size_t tab = (4u << t->host.lut_bits) + (t->host.has_long ? 256 + align_up(sizeof(LongTables), 16) : 0);
    const size_t warp_bytes = (size_t)(4 * segs + 2 + 2 * g.stage_rows) * kWarpRowBytes;
    if ((size_t)max_smem < tab + kTileStaticSmem + warp_bytes) return false;
    const size_t fit = ((size_t)max_smem - tab - kTileStaticSmem) / warp_bytes;
    g.n_warps = (uint32_t)std::min<size_t>(fit, want_warps ? want_warps : kTileMaxWarps);
    g.n_warps = std::min(g.n_warps, kTileMaxWarps);
    if (g.n_warps * 32u * (uint32_t)segs < kTileThreads) return false;   // the workspace holds one status word per 256 segments
    g.smem = tab + kWarpRowBytes + g.n_warps * warp_bytes;
    return true;
}

int g_tile_segs = 0, g_tile_warps = 0;   // hd_set_tuning: 0 = automatic

bool pick_tile_geom(const hd_table *t, int max_smem, TileGeom &g) {
    if (g_tile_segs && tile_geom(t, g_tile_segs, max_smem, (uint32_t)g_tile_warps, g)) return true;
    // Two segments per thread when at least 20 decode warps fit (the per-step work of a warp is amortised over twice
    // the symbols), else one segment per thread (short minimum code length: the worst-case staging of two does not
    // leave enough warps to hide the table latency).  A FIXED-LENGTH code of 1- or 2-bit codewords (every chunk decodes
    // to the same number of symbols) keeps one segment: the lanes' spans in the compacted block are then a constant
    // stride apart -- 32 words with two 2-bit segments, i.e. every lane of the warp starts in the same bank -- and the
    // smaller chunk halves the degree of that conflict (measured at redundancy 1.0: 0.96 ms against 1.44 ms per GiB).
    TileGeom g2, g1;
    const bool ok2 = tile_geom(t, 2, max_smem, (uint32_t)g_tile_warps, g2);
    const bool ok1 = tile_geom(t, 1, max_smem, (uint32_t)g_tile_warps, g1);
    const bool fixed_length = t->host.min_len == t->host.max_len && t->host.min_len <= 2;
    if (ok2 && ((g2.n_warps >= 20 && !fixed_length) || !ok1)) { g = g2; return true; }
    if (ok1) { g = g1; return true; }
    return false;
}

template <int SEGS, int RCAP, bool LONG>
hd_status launch_tiles(const TileParams &p, size_t smem, cudaStream_t s, bool after_sync) {
    auto kern = hd_decode_tiles_kernel<SEGS, RCAP, LONG>;
    HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t grid = (uint64_t)device_info().sm_count;   // persistent: one CTA per SM, all co-resident (the look-back spins)
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (after_sync && !(HD_WHATIF & 16384)) {
        // K4 right behind K1a on the stream: programmatic dependent launch.  K1a triggers at its start, so a K4 CTA is
        // placed on an SM as soon as that SM's K1a CTA has exited; it stages its tables and takes its first ticket while
        // the last K1a warps finish, and waits (cudaGridDependencySynchronize) before it reads the gap words.
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((p.n_warps + 1) * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        HD_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return HD_OK;
    }
    kern<<<(unsigned)grid, (p.n_warps + 1) * 32, smem, s>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

template <bool LONG>
hd_status launch_tiles_geom(const TileGeom &g, const TileParams &p, cudaStream_t s, bool after_sync) {
    switch (g.segs * 100 + g.rcap) {
        case 112: return launch_tiles<1, 12, LONG>(p, g.smem, s, after_sync);
        case 124: return launch_tiles<1, 24, LONG>(p, g.smem, s, after_sync);
        case 136: return launch_tiles<1, 36, LONG>(p, g.smem, s, after_sync);
        case 212: return launch_tiles<2, 12, LONG>(p, g.smem, s, after_sync);
        case 224: return launch_tiles<2, 24, LONG>(p, g.smem, s, after_sync);
        case 236: return launch_tiles<2, 36, LONG>(p, g.smem, s, after_sync);
        case 424: return launch_tiles<4, 24, LONG>(p, g.smem, s, after_sync);
    }
    return HD_ERR_INVALID;
}

// ---- fixed-length codes (hd_fixed.cuh) ---------------------------------------------------------------------------
// True if `t` is a complete code whose codewords all have the same length L in {1, 2, 4, 8}.
bool fixed_length_code(const hd_table *t, uint32_t &L) {
    const TableBlob &b = t->host;
    L = b.min_len;
    if (b.min_len != b.max_len || b.has_long) return false;   // (has_long also covers incomplete codes)
    return L == 1 || L == 2 || L == 4 || L == 8;
}

bool g_fixed_fast_path = true;   // hd_set_tuning(-1, 0) switches it off (tests decode these codes through K1a + K4 too)

template <int L>
hd_status launch_fixed(const FixedParams &q, cudaStream_t s) {
    const uint64_t total = q.n_segments * (128u / L);
    const uint64_t blocks = std::max<uint64_t>(1, div_up(div_up(std::min(total, q.n_symbols), 16), 512));   // 16 output bytes per thread
    const uint64_t cap = (uint64_t)device_info().sm_count * 4u;
    hd_decode_fixed_kernel<L><<<(unsigned)std::min(blocks, cap), 512, 0, s>>>(q);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

hd_status run_fixed_decode(const uint32_t *d_units, uint64_t n_units, uint64_t n_segments, const hd_table *table, uint32_t L,
                           uint8_t *d_out, uint64_t n_symbols, uint64_t *d_count, cudaStream_t s) {
    FixedParams q;
    q.units = d_units;
    q.n_units = n_units;
    q.n_segments = n_segments;
    q.out = d_out;
    q.n_symbols = n_symbols;
    q.count_out = d_count;
    // symbol of codeword c: first symbol of the primary LUT entry whose window starts with c
    const uint32_t *dec = reinterpret_cast<const uint32_t *>(table->host.bytes.data());
    const uint32_t lm = table->host.lut_bits;
    std::memset(q.sym, 0, sizeof(q.sym));
    for (uint32_t c = 0; c < (1u << L); ++c) q.sym[c] = (uint8_t)(dec[c << (lm - L)] >> 8);
    ProfScope prof(HD_PROFILE_DECODE_WRITE, s);
    switch (L) {
        case 1: return launch_fixed<1>(q, s);
        case 2: return launch_fixed<2>(q, s);
        case 4: return launch_fixed<4>(q, s);
        case 8: return launch_fixed<8>(q, s);
    }
    return HD_ERR_INVALID;
}

// The fast path applies to a whole stream that starts at bit 0 and writes to an aligned buffer; shards, chained pieces
// and streams entered at an offset keep the general kernels.
bool fixed_fast_ok(const hd_table *table, uint32_t first_bit, const uint8_t *d_out, uint32_t &L) {
    return g_fixed_fast_path && first_bit == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0 && fixed_length_code(table, L);
}

hd_status run_gap_decode(const uint32_t *d_units, uint64_t n_units, const uint32_t *d_gap_words,
                         uint64_t n_segments, uint32_t first_bit, const hd_table *table, uint8_t *d_out,
                         uint64_t n_symbols, const Workspace &w, uint64_t *d_count, cudaStream_t s,
                         uint64_t n_units_readable = 0, const uint32_t *entry_word = nullptr,
                         uint64_t *shard_info = nullptr, bool already_zeroed = false,
                         const uint64_t *d_off_in = nullptr, uint64_t *d_off_out = nullptr,
                         const uint32_t *unconverged = nullptr, bool after_sync = false) {
    uint32_t fixed_len = 0;
    if (!entry_word && !shard_info && !d_off_in && !d_off_out && !unconverged && fixed_fast_ok(table, first_bit, d_out, fixed_len))
        return run_fixed_decode(d_units, n_units_readable > n_units ? n_units_readable : n_units, n_segments, table, fixed_len,
                                d_out, n_symbols, d_count, s);
    TileGeom g;
    if (!pick_tile_geom(table, device_info().max_smem_optin, g)) return HD_ERR_INVALID;
    TileParams q;
    q.units = d_units;
    q.n_units = n_units_readable > n_units ? n_units_readable : n_units;
    q.gap_words = d_gap_words;
    q.n_segments = n_segments;
    q.first_bit = first_bit;
    q.entry_word = entry_word;
    q.shard_info = shard_info;
    q.unconverged = unconverged;
    q.table = table_dev(table);
    q.out = d_out;
    q.n_symbols = n_symbols;
    q.out_offset_in = d_off_in;
    q.out_offset_out = d_off_out;
    q.tile_status = w.tile_status;
    q.tile_counter = w.tile_counter;
    q.n_tiles = div_up(div_up(n_segments, (uint64_t)g.segs), (uint64_t)g.n_warps * 32);
    q.count_out = d_count;
    q.stage_rows = g.stage_rows;
    q.n_warps = g.n_warps;
    q.k_one = 1u;
    q.k_sel_len = kSelLen;
    q.k_sel_cnt8 = kSelCnt8;
    q.k_pair_slope = kPairSelSlope;
    if (!already_zeroed) HD_CUDA(cudaMemsetAsync(w.tile_counter, 0, w.zero_bytes, s));
    ProfScope prof(HD_PROFILE_DECODE_WRITE, s);
    return table->host.has_long ? launch_tiles_geom<true>(g, q, s, after_sync) : launch_tiles_geom<false>(g, q, s, after_sync);
}

hd_status check_common(const uint32_t *d_units, uint64_t n_units, const hd_table *table, const uint8_t *d_out,
                       uint64_t n_symbols, const void *ws) {
    if (!table || !table->d_blob) return HD_ERR_INVALID;
    if (n_units == 0) return n_symbols == 0 ? HD_OK : HD_ERR_INVALID;
    if (!d_units || (!d_out && n_symbols) || !ws) return HD_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(d_units) & 15u) return HD_ERR_INVALID;   // 128-bit segment loads
    if (reinterpret_cast<uintptr_t>(ws) & 63u) return HD_ERR_INVALID;
    return HD_OK;
}

// phases 1+2: build the gap array of a stream by self-synchronisation
struct SyncLaunch {
    SyncParams p;
    size_t smem_tables, smem_parse;
    bool exact, lng;
};

SyncLaunch make_sync(const uint32_t *d_units, uint64_t n_units, uint32_t first_bit, const hd_table *table,
                            uint32_t *d_gap_words, uint8_t *d_cur_entry, uint32_t *d_ctrl, uint64_t n_units_readable,
                            bool exact, uint32_t *d_sync_exit, unsigned long long *d_tile_map = nullptr,
                            uint8_t *d_tile_entry = nullptr) {
    SyncLaunch L;
    SyncParams &p = L.p;
    p.units = d_units;
    p.n_units = n_units_readable > n_units ? n_units_readable : n_units;
    p.n_segments = div_up(n_units, kSegmentUnits);
    p.n_chunks = div_up(p.n_segments, kSyncChunkSegs);
    p.n_tiles = std::max<uint64_t>(1, div_up(p.n_chunks - 1, kSyncTileChunks));   // tile t owns chunks 31t+1 .. 31t+31 (+ chunk 0)
    p.tile_exit = d_sync_exit;
    p.tile_map = d_tile_map;
    p.tile_entry = d_tile_entry;
    p.use_tile_entry = 0;
    p.k_one = 1u;
    p.first_bit = first_bit;
    p.table = table_dev(table);
    p.gap_words = d_gap_words;
    p.cur_entry = d_cur_entry;
    p.changed = d_ctrl;
    p.tile_counter = d_ctrl + 8;
    L.smem_tables = sync_smem_bytes(table);
    // (the decode path's K1a keeps a byte jump table, aligned to its size, instead of the u16 table: see the kernel)
    const size_t parse_tables = (exact || (HD_WHATIF & 2048)) ? L.smem_tables : (size_t)sync_jump8_tables_bytes(table->host.lut_bits);
    L.smem_parse = parse_tables + (size_t)kSyncWarps * kSyncTileBytes;
    L.exact = exact;
    L.lng = table->host.has_long != 0;
    return L;
}

template <bool EXACT, bool LONG>
hd_status launch_sync_parse(const SyncLaunch &L, cudaStream_t s) {
    auto kern = hd_sync_parse_kernel<EXACT, LONG>;
    HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_parse));
    uint64_t grid = (uint64_t)device_info().sm_count;   // persistent: one CTA per SM, warps pull tiles
    const uint64_t need = div_up(L.p.n_tiles, kSyncWarps);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, kSyncWarps * 32, L.smem_parse, s>>>(L.p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

template <bool EXACT, bool LONG>
hd_status launch_sync_walk(const SyncLaunch &L, uint32_t *changed, cudaStream_t s) {
    SyncParams p = L.p;
    p.changed = changed;
    auto kern = hd_sync_walk_kernel<EXACT, LONG>;
    const size_t smem = L.smem_tables + 128 * (kSyncChunkSegs + 2) * 16;   // + one chunk slot per thread
    HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::max<uint64_t>(1, div_up(div_up(p.n_chunks - 1, kSyncWalkChunks), 128));
    kern<<<grid, 128, smem, s>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// second-level fallback: transfer functions of all tiles (K1c), the chase through them, then K1a with exact tile entries
template <bool EXACT, bool LONG>
hd_status launch_sync_maps(const SyncLaunch &L, cudaStream_t s) {
    if (!L.p.tile_map || !L.p.tile_entry) return HD_ERR_UNRESOLVED;
    HD_CUDA(cudaMemsetAsync(L.p.tile_map, 0, L.p.n_tiles * sizeof(unsigned long long), s));
    auto kmap = hd_sync_map_kernel<EXACT, LONG>;
    auto kchase = hd_sync_chase_kernel<EXACT, LONG>;
    HD_CUDA(cudaFuncSetAttribute(kmap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_tables));
    HD_CUDA(cudaFuncSetAttribute(kchase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_tables));
    const unsigned grid = (unsigned)div_up(L.p.n_tiles * 16, 256);
    kmap<<<grid, 256, L.smem_tables, s>>>(L.p);
    kchase<<<1, 256, L.smem_tables, s>>>(L.p);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    SyncLaunch L2 = L;
    L2.p.use_tile_entry = 1;
    HD_CUDA(cudaMemsetAsync(L2.p.tile_counter, 0, 4, s));   // K1a's ticket counter
    return launch_sync_parse<EXACT, LONG>(L2, s);
}

#define HD_SYNC_DISPATCH(fn, L, ...)                                                               \
    ((L).exact ? ((L).lng ? fn<true, true>((L), __VA_ARGS__) : fn<true, false>((L), __VA_ARGS__))   \
               : ((L).lng ? fn<false, true>((L), __VA_ARGS__) : fn<false, false>((L), __VA_ARGS__)))

// K1a alone: every warp tile publishes the offset at which its parse leaves it, and its warp later checks the tile's
// first seam against the exit the predecessor published (one tile later, when that exit has long been there: no wait
// and no second launch).  changed[0] counts the tiles whose exit moved AFTER they had published it (a seam repair that
// ran through the whole tile) and the waits that timed out: zero means every seam was checked against an exit that
// never changed afterwards, i.e. the fixed point is reached.  d_ctrl + the published exits must be zeroed
// (Workspace::sync_zero_bytes) before.
hd_status run_sync(const SyncLaunch &L, cudaStream_t s) {
    ProfScope prof(HD_PROFILE_SYNC_CHUNKS, s);
    return HD_SYNC_DISPATCH(launch_sync_parse, L, s);
}

// K1b rounds until no tile exit changes (host loop of the reference: cuhd_gpu_decoder.cu:461-495); only reached for
// codes that fail to synchronise inside a whole warp tile (64 Kbit).
hd_status finish_sync(const SyncLaunch &L, cudaStream_t s, bool *extra_rounds) {
    *extra_rounds = false;
    if (L.p.n_chunks <= 1) return HD_OK;
    uint32_t h_changed = 0;
    HD_CUDA(cudaMemcpyAsync(&h_changed, L.p.changed, 4, cudaMemcpyDeviceToHost, s));
    HD_CUDA(cudaStreamSynchronize(s));
    if (h_changed == 0) return HD_OK;
    *extra_rounds = true;
    uint32_t *changed = L.p.changed + 2;
    // A parse that fails to merge inside one tile usually merges in the next: a few K1b rounds settle it.  A code that
    // keeps failing (it does not self-synchronise at all) would need one round per tile: the transfer-function pass
    // resolves every tile at once instead.
    const uint64_t max_rounds = 3;
    for (uint64_t round = 0; round <= max_rounds; ++round) {
        HD_CUDA(cudaMemsetAsync(changed, 0, 4, s));
        ProfScope prof(HD_PROFILE_SYNC_SEAMS, s);
        hd_status st = HD_SYNC_DISPATCH(launch_sync_walk, L, changed, s);
        if (st != HD_OK) return st;
        HD_CUDA(cudaMemcpyAsync(&h_changed, changed, 4, cudaMemcpyDeviceToHost, s));
        HD_CUDA(cudaStreamSynchronize(s));
        if (h_changed == 0) return HD_OK;
    }
    {
        ProfScope prof(HD_PROFILE_SYNC_SEAMS, s);
        hd_status st = L.exact ? (L.lng ? launch_sync_maps<true, true>(L, s) : launch_sync_maps<true, false>(L, s))
                               : (L.lng ? launch_sync_maps<false, true>(L, s) : launch_sync_maps<false, false>(L, s));
        if (st != HD_OK) return st;
    }
    HD_CUDA(cudaStreamSynchronize(s));
    return HD_OK;
}

// the pinned word a synchronous self-sync decode reads its convergence flag into: one per host thread, made once
// (hd_warmup makes it ahead of the first decode)
uint32_t *thread_flag() {
    static thread_local uint32_t *h_flag = nullptr;
    if (!h_flag && cudaHostAlloc(reinterpret_cast<void **>(&h_flag), 64, cudaHostAllocDefault) != cudaSuccess) h_flag = nullptr;
    return h_flag;
}

}  // namespace

extern "C" {

const char *hd_version(void) { return "hd_b200 0.1 (sm_100a)"; }

const char *hd_status_string(hd_status s) {
    switch (s) {
        case HD_OK: return "ok";
        case HD_ERR_INVALID: return "invalid argument";
        case HD_ERR_CUDA: return "CUDA error";
        case HD_ERR_WORKSPACE: return "workspace too small";
        case HD_ERR_FORMAT: return "malformed input";
        case HD_ERR_UNRESOLVED: return "self-synchronisation did not converge";
    }
    return "unknown";
}

const char *hd_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda); }

// Creates the CUDA context of the current device and loads the decode kernels (CUDA loads a kernel's code lazily, at
// its first launch: tens of milliseconds that would otherwise land inside the first timed decode of a process).
hd_status hd_warmup(void) {
    HD_CUDA(cudaFree(nullptr));
    (void)device_info();
    if (!thread_flag()) { g_last_cuda = cudaGetLastError(); return HD_ERR_CUDA; }   // (a pinned allocation costs milliseconds)
    cudaFuncAttributes a;
#define HD_TOUCH(k) HD_CUDA(cudaFuncGetAttributes(&a, k))
#define HD_TOUCH_TILES(L)                                                                                  \
    HD_TOUCH((hd_decode_tiles_kernel<1, 12, L>)); HD_TOUCH((hd_decode_tiles_kernel<1, 24, L>));            \
    HD_TOUCH((hd_decode_tiles_kernel<1, 36, L>)); HD_TOUCH((hd_decode_tiles_kernel<2, 12, L>));            \
    HD_TOUCH((hd_decode_tiles_kernel<2, 24, L>)); HD_TOUCH((hd_decode_tiles_kernel<2, 36, L>)); HD_TOUCH((hd_decode_tiles_kernel<4, 24, L>));
    HD_TOUCH_TILES(false)
    HD_TOUCH_TILES(true)
    HD_TOUCH((hd_sync_parse_kernel<false, false>)); HD_TOUCH((hd_sync_parse_kernel<false, true>));
    HD_TOUCH((hd_sync_parse_kernel<true, false>)); HD_TOUCH((hd_sync_parse_kernel<true, true>));
    HD_TOUCH((hd_sync_walk_kernel<false, false>)); HD_TOUCH((hd_sync_walk_kernel<false, true>));
    HD_TOUCH((hd_sync_walk_kernel<true, false>)); HD_TOUCH((hd_sync_walk_kernel<true, true>));
    HD_TOUCH((hd_decode_fixed_kernel<1>)); HD_TOUCH((hd_decode_fixed_kernel<2>));
    HD_TOUCH((hd_decode_fixed_kernel<4>)); HD_TOUCH((hd_decode_fixed_kernel<8>));
#undef HD_TOUCH_TILES
#undef HD_TOUCH
    (void)a;
    return HD_OK;
}

// ---------------------------------------------------------------- tables
static hd_status upload_table(const HostCode &c, void *stream, hd_table **out) {
    hd_table *t = new (std::nothrow) hd_table();
    if (!t) return HD_ERR_INVALID;
    build_table_blob(c, t->host);
    cudaError_t e = cudaGetDevice(&t->device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&t->d_blob), t->host.total_bytes);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(t->d_blob, t->host.bytes.data(), t->host.total_bytes, cudaMemcpyHostToDevice,
                            static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        g_last_cuda = e;
        (void)cudaGetLastError();
        if (t->d_blob) cudaFree(t->d_blob);
        delete t;
        return HD_ERR_CUDA;
    }
    *out = t;
    return HD_OK;
}

hd_status hd_table_from_flat_lut(const hd_lut_entry *lut, uint32_t lut_bits, void *stream, hd_table **out) {
    if (!out) return HD_ERR_INVALID;
    HostCode c;
    if (!code_from_flat_lut(lut, lut_bits, c)) return HD_ERR_FORMAT;
    return upload_table(c, stream, out);
}

hd_status hd_table_from_lengths(const hd_sym_len *pairs, uint32_t n_pairs, void *stream, hd_table **out) {
    if (!out) return HD_ERR_INVALID;
    HostCode c;
    if (!code_from_pairs(pairs, n_pairs, c)) return HD_ERR_FORMAT;
    return upload_table(c, stream, out);
}

// K0: the whole blob is built by a kernel from a DEVICE-resident {symbol,length} list; only a 20-byte summary comes
// back (the host needs lut_bits / min_len / max_len to pick kernel variants and shared-memory geometry).
hd_status hd_table_from_lengths_device(const hd_sym_len *d_pairs, uint32_t n_pairs, void *stream, hd_table **out) {
    if (!out || !d_pairs || n_pairs == 0 || n_pairs > 256) return HD_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    hd_table *t = new (std::nothrow) hd_table();
    if (!t) return HD_ERR_INVALID;
    TableSummary *d_sum = nullptr;
    TableSummary h_sum = {0, 0, 0, 0, 0};
    cudaError_t e = cudaGetDevice(&t->device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&t->d_blob), kTableBlobMaxBytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&d_sum), sizeof(TableSummary));
    if (e == cudaSuccess) e = cudaMemsetAsync(t->d_blob, 0, kTableBlobMaxBytes, s);
    if (e == cudaSuccess) {
        hd_build_tables_kernel<<<1, 1024, 0, s>>>(d_pairs, n_pairs, t->d_blob, d_sum);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_sum, d_sum, sizeof(h_sum), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (d_sum) cudaFree(d_sum);
    if (e != cudaSuccess || !h_sum.ok) {
        if (e != cudaSuccess) { g_last_cuda = e; (void)cudaGetLastError(); }
        if (t->d_blob) cudaFree(t->d_blob);
        delete t;
        return e != cudaSuccess ? HD_ERR_CUDA : HD_ERR_FORMAT;
    }
    TableBlob &b = t->host;
    const uint32_t n = 1u << h_sum.lut_bits;
    b.lut_bits = h_sum.lut_bits;
    b.min_len = h_sum.min_len;
    b.max_len = h_sum.max_len;
    b.has_long = h_sum.has_long;
    b.off_sync = n * 4;
    b.off_len = b.off_sync + n * 2;
    b.off_order = b.off_len + 256;
    b.off_long = b.off_order + 256;
    b.total_bytes = (uint32_t)align_up(b.off_long + sizeof(LongTables), 16);
    *out = t;
    return HD_OK;
}

// device copy of a table's blob (tests compare the device-built blob with the host-built one)
hd_status hd_table_blob(const hd_table *t, void *dst, size_t cap, size_t *size) {
    if (!t || !t->d_blob) return HD_ERR_INVALID;
    if (size) *size = t->host.total_bytes;
    if (!dst) return HD_OK;
    if (cap < t->host.total_bytes) return HD_ERR_INVALID;
    HD_CUDA(cudaMemcpy(dst, t->d_blob, t->host.total_bytes, cudaMemcpyDeviceToHost));
    return HD_OK;
}

void hd_table_free(hd_table *t) {
    if (!t) return;
    if (t->d_blob) cudaFree(t->d_blob);
    delete t;
}

uint32_t hd_table_max_length(const hd_table *t) { return t ? t->host.max_len : 0; }
uint32_t hd_table_min_length(const hd_table *t) { return t ? t->host.min_len : 0; }

// ------------------------------------------------------------- workspace
size_t hd_workspace_bytes_gap(uint64_t n_units) { return ws_bytes(n_units, false); }
size_t hd_workspace_bytes_selfsync(uint64_t n_units) { return ws_bytes(n_units, true); }

// ---------------------------------------------------------------- decode
hd_status hd_decode_gap_ex(const uint32_t *d_units, uint64_t n_units, const uint32_t *d_gap_words,
                           uint64_t n_segments, uint32_t first_bit, const hd_table *table, uint8_t *d_out,
                           uint64_t n_symbols, void *d_workspace, size_t workspace_bytes, uint64_t *d_count,
                           void *stream) {
    hd_status st = check_common(d_units, n_units, table, d_out, n_symbols, d_workspace);
    if (st != HD_OK) return st;
    if (n_units == 0) return HD_OK;
    // n_units = units that may be READ; a shard of a longer stream passes fewer segments than that so that its
    // last codeword can end in the right halo (the stop rule is "codeword STARTS inside the segment")
    if (n_segments == 0 || n_segments > div_up(n_units, kSegmentUnits) || first_bit > 15) return HD_ERR_INVALID;
    if (n_segments > 1 && !d_gap_words) return HD_ERR_INVALID;
    if (workspace_bytes < ws_bytes(n_units, false)) return HD_ERR_WORKSPACE;
    Workspace w = carve(d_workspace, n_units, false);
    // The kernel reads gap element n_segments - 1 (the offset at which the parse leaves the last segment; only the
    // multi-GPU seam check uses its value).  A one-segment stream may come without a gap array: point the kernel at a
    // word of the control block that the per-launch memset zeroes.
    if (!d_gap_words) d_gap_words = w.tile_counter + 8;
    return run_gap_decode(d_units, n_units, d_gap_words, n_segments, first_bit, table, d_out, n_symbols, w,
                          d_count, static_cast<cudaStream_t>(stream));
}

// One shard of a gap-array stream (multi-GPU).  d_gap_words[-1] must be readable when d_entry_from_prev is set: its top
// nibble is the offset at which the shard's first segment starts (decoder.cu:505-507).
hd_status hd_decode_gap_shard(const uint32_t *d_units, uint64_t n_units_readable, const uint32_t *d_gap_words,
                              uint64_t n_segments, int entry_from_prev_word, const hd_table *table, uint8_t *d_out,
                              uint64_t out_cap, void *d_workspace, size_t workspace_bytes, uint64_t *d_shard_info,
                              void *stream) {
    hd_status st = check_common(d_units, n_units_readable, table, d_out, out_cap, d_workspace);
    if (st != HD_OK) return st;
    if (n_units_readable == 0 || n_segments == 0 || n_segments > div_up(n_units_readable, kSegmentUnits) || !d_gap_words)
        return HD_ERR_INVALID;
    if (workspace_bytes < ws_bytes(n_units_readable, false)) return HD_ERR_WORKSPACE;
    Workspace w = carve(d_workspace, n_units_readable, false);
    return run_gap_decode(d_units, n_units_readable, d_gap_words, n_segments, 0, table, d_out, out_cap, w, nullptr,
                          static_cast<cudaStream_t>(stream), 0, entry_from_prev_word ? d_gap_words - 1 : nullptr,
                          d_shard_info);
}

hd_status hd_decode_gap(const uint32_t *d_units, uint64_t n_units, const uint32_t *d_gap_words,
                        uint64_t n_segments, const hd_table *table, uint8_t *d_out, uint64_t n_symbols,
                        void *d_workspace, size_t workspace_bytes, void *stream) {
    return hd_decode_gap_ex(d_units, n_units, d_gap_words, n_segments, 0, table, d_out, n_symbols, d_workspace,
                            workspace_bytes, nullptr, stream);
}

hd_status hd_selfsync_gap_array(const uint32_t *d_units, uint64_t n_units, uint32_t first_bit,
                                const hd_table *table, uint32_t *d_gap_words_out, void *d_workspace,
                                size_t workspace_bytes, void *stream) {
    hd_status st = check_common(d_units, n_units, table, nullptr, 0, d_workspace);
    if (st != HD_OK) return st;
    if (n_units == 0) return HD_OK;
    if (!d_gap_words_out || first_bit > 15) return HD_ERR_INVALID;
    if (workspace_bytes < ws_bytes(n_units, true)) return HD_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace w = carve(d_workspace, n_units, true);
    HD_CUDA(cudaMemsetAsync(w.changed, 0, w.sync_zero_bytes, s));
    // exact landings: this entry returns a gap array in the .huff sense (first codeword of every segment)
    const SyncLaunch L = make_sync(d_units, n_units, first_bit, table, w.gap_words, w.blk_entry, w.changed, 0, true, w.sync_exit, w.tile_map, w.tile_entry);
    st = run_sync(L, s);
    if (st != HD_OK) return st;
    bool extra = false;
    st = finish_sync(L, s, &extra);
    if (st != HD_OK) return st;
    const uint64_t n_seg = div_up(n_units, kSegmentUnits);
    // chunk words are exactly the .huff gap words (8 nibbles each); the last
    // word may describe segments past the end, which is harmless
    HD_CUDA(cudaMemcpyAsync(d_gap_words_out, w.gap_words, div_up(n_seg, 8) * 4, cudaMemcpyDeviceToDevice, s));
    return HD_OK;
}

// Asynchronous half of the self-sync decode: one memset, K1a, K1b, K4 and the copy of the convergence counter into
// *h_unconverged (pinned host memory for a truly asynchronous copy), all on `stream`; returns without synchronising.
hd_status hd_decode_selfsync_enqueue(const uint32_t *d_units, uint64_t n_units, uint32_t first_bit,
                                     const hd_table *table, uint8_t *d_out, uint64_t n_symbols, void *d_workspace,
                                     size_t workspace_bytes, uint64_t *d_count, uint32_t *h_unconverged,
                                     void *stream) {
    hd_status st = check_common(d_units, n_units, table, d_out, n_symbols, d_workspace);
    if (st != HD_OK) return st;
    if (h_unconverged) *h_unconverged = 0;
    if (n_units == 0) return HD_OK;
    if (first_bit > 15 || !h_unconverged) return HD_ERR_INVALID;
    if (workspace_bytes < ws_bytes(n_units, true)) return HD_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace w = carve(d_workspace, n_units, true);
    const uint64_t n_seg = div_up(n_units, kSegmentUnits);
    // a fixed-length code needs no synchronisation: every segment starts at a codeword boundary (hd_fixed.cuh)
    uint32_t fixed_len = 0;
    if (fixed_fast_ok(table, first_bit, d_out, fixed_len))
        return run_fixed_decode(d_units, n_units, n_seg, table, fixed_len, d_out, n_symbols, d_count, s);
    // K1a + K1b, then K4 speculatively; the convergence flag is read once at the end (the reference reads a flag
    // array after every phase-2 launch, cuhd_gpu_decoder.cu:461-495).  One memset covers K4's ticket counter and
    // tile statuses and, right behind them, K1's control block.
    HD_CUDA(cudaMemsetAsync(w.tile_counter, 0, w.zero_bytes + w.sync_zero_bytes, s));
    const SyncLaunch L = make_sync(d_units, n_units, first_bit, table, w.gap_words, w.blk_entry, w.changed, 0, false, w.sync_exit, w.tile_map, w.tile_entry);
    st = run_sync(L, s);
    if (st != HD_OK) return st;
    st = run_gap_decode(d_units, n_units, w.gap_words, n_seg, first_bit, table, d_out, n_symbols, w, d_count, s, 0, nullptr,
                        nullptr, true, nullptr, nullptr, nullptr, true);
    if (st != HD_OK) return st;
    if (L.p.n_chunks > 1) HD_CUDA(cudaMemcpyAsync(h_unconverged, w.changed, 4, cudaMemcpyDeviceToHost, s));
    return HD_OK;
}

// Synchronous half: waits for the stream; if the blind K1b round moved a group exit (a parse that did not
// synchronise within 64 Kbit), runs further rounds to the fixed point and decodes again.
hd_status hd_decode_selfsync_finish(const uint32_t *d_units, uint64_t n_units, uint32_t first_bit,
                                    const hd_table *table, uint8_t *d_out, uint64_t n_symbols, void *d_workspace,
                                    size_t workspace_bytes, uint64_t *d_count, const uint32_t *h_unconverged,
                                    void *stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HD_CUDA(cudaStreamSynchronize(s));
    if (n_units == 0 || !h_unconverged || *h_unconverged == 0) return HD_OK;
    hd_status st = check_common(d_units, n_units, table, d_out, n_symbols, d_workspace);
    if (st != HD_OK) return st;
    if (workspace_bytes < ws_bytes(n_units, true)) return HD_ERR_WORKSPACE;
    Workspace w = carve(d_workspace, n_units, true);
    const SyncLaunch L = make_sync(d_units, n_units, first_bit, table, w.gap_words, w.blk_entry, w.changed, 0, false, w.sync_exit, w.tile_map, w.tile_entry);
    bool extra = false;
    st = finish_sync(L, s, &extra);
    if (st != HD_OK) return st;
    st = run_gap_decode(d_units, n_units, w.gap_words, div_up(n_units, kSegmentUnits), first_bit, table, d_out, n_symbols,
                        w, d_count, s);
    if (st != HD_OK) return st;
    HD_CUDA(cudaStreamSynchronize(s));
    return HD_OK;
}

hd_status hd_decode_selfsync_ex(const uint32_t *d_units, uint64_t n_units, uint32_t first_bit,
                                const hd_table *table, uint8_t *d_out, uint64_t n_symbols, void *d_workspace,
                                size_t workspace_bytes, uint64_t *d_count, void *stream) {
    uint32_t *h_flag = thread_flag();   // pinned, one per host thread
    if (!h_flag) { g_last_cuda = cudaGetLastError(); return HD_ERR_CUDA; }
    hd_status st = hd_decode_selfsync_enqueue(d_units, n_units, first_bit, table, d_out, n_symbols, d_workspace,
                                              workspace_bytes, d_count, h_flag, stream);
    if (st != HD_OK) return st;
    return hd_decode_selfsync_finish(d_units, n_units, first_bit, table, d_out, n_symbols, d_workspace, workspace_bytes,
                                     d_count, h_flag, stream);
}

// One shard of a longer stream (multi-GPU): [halo | shard | right halo].  The
// halo is parsed speculatively from bit 0 so that the parse is synchronised
// when it enters the shard; the entry offset it reaches and the offset at which
// the shard's parse leaves are reported so that neighbours can be checked
// against each other (entry of rank g == exit of rank g-1).
static hd_status shard_args(const uint32_t *d_units, uint64_t n_units_halo, uint64_t n_units_shard,
                            uint64_t &n_units_readable, uint32_t first_bit, const hd_table *table, uint8_t *d_out,
                            uint64_t out_cap, void *d_workspace, size_t workspace_bytes) {
    const uint64_t n_units = n_units_halo + n_units_shard;
    hd_status st = check_common(d_units, n_units, table, d_out, out_cap, d_workspace);
    if (st != HD_OK) return st;
    if (n_units_shard == 0 || first_bit > 15) return HD_ERR_INVALID;
    if (n_units_halo % (kSegmentUnits * 8)) return HD_ERR_INVALID;   // gap words (8 segments each) must line up
    if (n_units_readable < n_units) n_units_readable = n_units;
    if (workspace_bytes < ws_bytes(n_units, true)) return HD_ERR_WORKSPACE;
    return HD_OK;
}

static hd_status shard_decode(const uint32_t *d_units, uint64_t n_units_halo, uint64_t n_units_shard,
                              uint64_t n_units_readable, uint32_t first_bit, const hd_table *table, uint8_t *d_out,
                              uint64_t out_cap, const Workspace &w, uint64_t *d_shard_info, cudaStream_t s,
                              bool already_zeroed, const uint64_t *d_off_in = nullptr, uint64_t *d_off_out = nullptr,
                              bool speculative = false) {
    const uint64_t halo_words = n_units_halo / (kSegmentUnits * 8);
    const uint32_t *gaps = w.gap_words + halo_words;
    // speculative = decoded before the convergence counter of the synchronisation has been looked at: the counter is
    // folded into the published triple so that every rank of a sharded decode learns that this shard is decoded again
    return run_gap_decode(d_units + n_units_halo, n_units_shard, gaps, div_up(n_units_shard, kSegmentUnits), first_bit,
                          table, d_out, out_cap, w, nullptr, s, n_units_readable - n_units_halo,
                          n_units_halo ? gaps - 1 : nullptr, d_shard_info, already_zeroed, d_off_in, d_off_out,
                          speculative ? w.changed : nullptr);
}

// asynchronous half (see hd_decode_selfsync_enqueue); d_off_in / d_off_out chain the output position of consecutive
// pieces on the device (host-buffer pipeline)
static hd_status shard_enqueue_impl(const uint32_t *d_units, uint64_t n_units_halo, uint64_t n_units_shard,
                                    uint64_t n_units_readable, uint32_t first_bit, const hd_table *table,
                                    uint8_t *d_out, uint64_t out_cap, void *d_workspace, size_t workspace_bytes,
                                    uint64_t *d_shard_info, uint32_t *h_unconverged, void *stream,
                                    const uint64_t *d_off_in, uint64_t *d_off_out) {
    hd_status st = shard_args(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit, table, d_out, out_cap,
                              d_workspace, workspace_bytes);
    if (st != HD_OK) return st;
    if (!h_unconverged) return HD_ERR_INVALID;
    *h_unconverged = 0;
    const uint64_t n_units = n_units_halo + n_units_shard;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace w = carve(d_workspace, n_units, true);
    HD_CUDA(cudaMemsetAsync(w.tile_counter, 0, w.zero_bytes + w.sync_zero_bytes, s));
    const SyncLaunch L = make_sync(d_units, n_units, n_units_halo ? 0u : first_bit, table, w.gap_words, w.blk_entry,
                                   w.changed, n_units_readable, false, w.sync_exit, w.tile_map, w.tile_entry);
    st = run_sync(L, s);
    if (st != HD_OK) return st;
    st = shard_decode(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit, table, d_out, out_cap, w,
                      d_shard_info, s, true, d_off_in, d_off_out, true);
    if (st != HD_OK) return st;
    if (L.p.n_chunks > 1) HD_CUDA(cudaMemcpyAsync(h_unconverged, w.changed, 4, cudaMemcpyDeviceToHost, s));
    return HD_OK;
}

hd_status hd_decode_selfsync_shard_enqueue(const uint32_t *d_units, uint64_t n_units_halo, uint64_t n_units_shard,
                                           uint64_t n_units_readable, uint32_t first_bit, const hd_table *table,
                                           uint8_t *d_out, uint64_t out_cap, void *d_workspace,
                                           size_t workspace_bytes, uint64_t *d_shard_info, uint32_t *h_unconverged,
                                           void *stream) {
    return shard_enqueue_impl(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit, table, d_out, out_cap,
                              d_workspace, workspace_bytes, d_shard_info, h_unconverged, stream, nullptr, nullptr);
}

// synchronous half (see hd_decode_selfsync_finish)
hd_status hd_decode_selfsync_shard_finish(const uint32_t *d_units, uint64_t n_units_halo, uint64_t n_units_shard,
                                          uint64_t n_units_readable, uint32_t first_bit, const hd_table *table,
                                          uint8_t *d_out, uint64_t out_cap, void *d_workspace,
                                          size_t workspace_bytes, uint64_t *d_shard_info,
                                          const uint32_t *h_unconverged, void *stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HD_CUDA(cudaStreamSynchronize(s));
    if (!h_unconverged || *h_unconverged == 0) return HD_OK;
    hd_status st = shard_args(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit, table, d_out, out_cap,
                              d_workspace, workspace_bytes);
    if (st != HD_OK) return st;
    const uint64_t n_units = n_units_halo + n_units_shard;
    Workspace w = carve(d_workspace, n_units, true);
    const SyncLaunch L = make_sync(d_units, n_units, n_units_halo ? 0u : first_bit, table, w.gap_words, w.blk_entry,
                                   w.changed, n_units_readable, false, w.sync_exit, w.tile_map, w.tile_entry);
    bool extra = false;
    st = finish_sync(L, s, &extra);
    if (st != HD_OK) return st;
    st = shard_decode(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit, table, d_out, out_cap, w,
                      d_shard_info, s, false);
    if (st != HD_OK) return st;
    HD_CUDA(cudaStreamSynchronize(s));
    return HD_OK;
}

hd_status hd_decode_selfsync_shard(const uint32_t *d_units, uint64_t n_units_halo, uint64_t n_units_shard,
                                   uint64_t n_units_readable, uint32_t first_bit, const hd_table *table,
                                   uint8_t *d_out, uint64_t out_cap, void *d_workspace, size_t workspace_bytes,
                                   uint64_t *d_shard_info, void *stream) {
    static thread_local uint32_t *h_flag = nullptr;   // pinned, one per host thread
    if (!h_flag) HD_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&h_flag), 64, cudaHostAllocDefault));
    hd_status st = hd_decode_selfsync_shard_enqueue(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit,
                                                    table, d_out, out_cap, d_workspace, workspace_bytes, d_shard_info,
                                                    h_flag, stream);
    if (st != HD_OK) return st;
    return hd_decode_selfsync_shard_finish(d_units, n_units_halo, n_units_shard, n_units_readable, first_bit, table,
                                           d_out, out_cap, d_workspace, workspace_bytes, d_shard_info, h_flag, stream);
}

hd_status hd_decode_selfsync(const uint32_t *d_units, uint64_t n_units, const hd_table *table, uint8_t *d_out,
                             uint64_t n_symbols, void *d_workspace, size_t workspace_bytes, void *stream) {
    return hd_decode_selfsync_ex(d_units, n_units, 0, table, d_out, n_symbols, d_workspace, workspace_bytes,
                                 nullptr, stream);
}

// ------------------------------------------------------- device-side encoder (input side)
size_t hd_workspace_bytes_encode(uint64_t n) {
    const uint64_t n_tiles = std::max<uint64_t>(1, div_up(n, kEncTileSyms));
    return 64 + 1024 + align_up(n_tiles * 8 * kStatusStride, 64);
}

hd_status hd_histogram_device(const uint8_t *d_data, uint64_t n, uint64_t *d_hist, void *stream) {
    if (!d_data || !d_hist || n == 0) return HD_ERR_INVALID;
    if (reinterpret_cast<uintptr_t>(d_data) & 15u) return HD_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HD_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint64_t), s));
    const DeviceInfo di = device_info();
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)di.sm_count * 8, div_up(n, 256 * 16));
    hd_histogram_kernel<<<grid ? grid : 1, 256, 0, s>>>(d_data, n, reinterpret_cast<unsigned long long *>(d_hist));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// shared tail of the two encode entries: the code table {code | len << 24} x 256 is already at ws + 64
static hd_status encode_launch(const uint8_t *d_data, uint64_t n, uint32_t *d_units, uint64_t cap_units, uint32_t *d_gap_words,
                               uint64_t cap_gap_words, uint8_t *ws, uint64_t *d_bits_out, cudaStream_t s) {
    EncodeParams p;
    p.data = d_data;
    p.n = n;
    p.tile_counter = reinterpret_cast<uint32_t *>(ws);
    p.code_tab = reinterpret_cast<uint32_t *>(ws + 64);
    p.tile_status = reinterpret_cast<uint64_t *>(ws + 64 + 1024);
    p.units = d_units;
    p.cap_units = cap_units;
    p.gap_words = d_gap_words;
    p.n_tiles = div_up(n, kEncTileSyms);
    p.bits_out = d_bits_out;
    HD_CUDA(cudaMemsetAsync(d_units, 0, cap_units * 4, s));
    if (d_gap_words) HD_CUDA(cudaMemsetAsync(d_gap_words, 0, cap_gap_words * 4, s));
    const DeviceInfo di = device_info();
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)di.sm_count * 4, p.n_tiles);   // all resident: look-back spins
    hd_encode_kernel<<<grid, kEncThreads, 0, s>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

hd_status hd_encode_device(const uint8_t *d_data, uint64_t n, const hd_sym_len *pairs, uint32_t n_pairs,
                           uint32_t *d_units, uint64_t cap_units, uint32_t *d_gap_words, uint64_t cap_gap_words,
                           void *d_workspace, size_t workspace_bytes, uint64_t *d_bits_out, void *stream) {
    if (!d_data || !pairs || !d_units || !d_workspace || n == 0 || cap_units == 0) return HD_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) || (reinterpret_cast<uintptr_t>(d_workspace) & 63u)) return HD_ERR_INVALID;
    if (workspace_bytes < hd_workspace_bytes_encode(n)) return HD_ERR_WORKSPACE;
    HostCode c;
    if (!code_from_pairs(pairs, n_pairs, c)) return HD_ERR_FORMAT;
    uint32_t tab[256];
    for (int i = 0; i < 256; ++i) tab[i] = c.len[i] ? (c.code[i] | ((uint32_t)c.len[i] << 24)) : 0u;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t *ws = static_cast<uint8_t *>(d_workspace);
    HD_CUDA(cudaMemsetAsync(ws, 0, hd_workspace_bytes_encode(n), s));
    HD_CUDA(cudaMemcpyAsync(ws + 64, tab, sizeof(tab), cudaMemcpyHostToDevice, s));
    HD_CUDA(cudaStreamSynchronize(s));   // `tab` lives on this stack frame
    return encode_launch(d_data, n, d_units, cap_units, d_gap_words, cap_gap_words, ws, d_bits_out, s);
}

// The whole producer without the host: histogram (already in d_hist256) -> length-limited canonical code built by a
// kernel (E0) -> pack + gap array.  d_pairs_out (256 pairs) / d_n_pairs_out receive the .huff header list; nothing is
// read back by this call (the caller copies the pairs and the bit count when it writes the container).
size_t hd_workspace_bytes_code(void) { return sizeof(CodeScratch) + 64; }

hd_status hd_build_code_device(const uint64_t *d_hist256, uint32_t max_len, hd_sym_len *d_pairs_out, uint32_t *d_n_pairs_out,
                               uint32_t *d_code_tab256, void *d_scratch, size_t scratch_bytes, void *stream) {
    if (!d_hist256 || !d_pairs_out || !d_n_pairs_out || !d_code_tab256 || !d_scratch || max_len == 0 || max_len > kMaxCodeLen)
        return HD_ERR_INVALID;
    if (scratch_bytes < hd_workspace_bytes_code()) return HD_ERR_WORKSPACE;
    uintptr_t a = (reinterpret_cast<uintptr_t>(d_scratch) + 63u) & ~(uintptr_t)63u;
    hd_build_code_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(d_hist256), max_len, d_pairs_out, d_n_pairs_out, d_code_tab256,
        reinterpret_cast<CodeScratch *>(a));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

hd_status hd_encode_device_auto(const uint8_t *d_data, uint64_t n, uint32_t max_len, hd_sym_len *d_pairs_out,
                                uint32_t *d_n_pairs_out, uint32_t *d_units, uint64_t cap_units, uint32_t *d_gap_words,
                                uint64_t cap_gap_words, void *d_workspace, size_t workspace_bytes, uint64_t *d_bits_out,
                                void *stream) {
    if (!d_data || !d_pairs_out || !d_n_pairs_out || !d_units || !d_workspace || n == 0 || cap_units == 0) return HD_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) || (reinterpret_cast<uintptr_t>(d_workspace) & 63u)) return HD_ERR_INVALID;
    const size_t enc_bytes = align_up(hd_workspace_bytes_encode(n), 64);
    if (workspace_bytes < enc_bytes + 256 * sizeof(uint64_t) + hd_workspace_bytes_code()) return HD_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t *ws = static_cast<uint8_t *>(d_workspace);
    uint64_t *d_hist = reinterpret_cast<uint64_t *>(ws + enc_bytes);
    void *d_scratch = ws + enc_bytes + 256 * sizeof(uint64_t);
    HD_CUDA(cudaMemsetAsync(ws, 0, hd_workspace_bytes_encode(n), s));
    hd_status st = hd_histogram_device(d_data, n, d_hist, stream);
    if (st != HD_OK) return st;
    st = hd_build_code_device(d_hist, max_len, d_pairs_out, d_n_pairs_out, reinterpret_cast<uint32_t *>(ws + 64), d_scratch,
                              hd_workspace_bytes_code(), stream);
    if (st != HD_OK) return st;
    return encode_launch(d_data, n, d_units, cap_units, d_gap_words, cap_gap_words, ws, d_bits_out, s);
}

// ------------------------------------------------------ multi-GPU mailbox (peer-memory exchange over NVLink)
struct hd_mailbox {
    uint32_t world = 0, rank = 0;
    int device = 0;
    unsigned long long *local = nullptr;
    MailboxPeers peers = {};
    bool opened[kMailboxMaxWorld] = {};
    uint32_t *d_timeout = nullptr;
};

void hd_mailbox_destroy(hd_mailbox *m) {
    if (!m) return;
    for (uint32_t p = 0; p < m->world; ++p)
        if (m->opened[p]) cudaIpcCloseMemHandle(m->peers.buf[p]);
    if (m->local) cudaFree(m->local);
    if (m->d_timeout) cudaFree(m->d_timeout);
    delete m;
}

hd_status hd_mailbox_create(uint32_t world, uint32_t rank, hd_mailbox **out) {
    if (!out || world == 0 || world > kMailboxMaxWorld || rank >= world) return HD_ERR_INVALID;
    hd_mailbox *m = new (std::nothrow) hd_mailbox();
    if (!m) return HD_ERR_INVALID;
    m->world = world;
    m->rank = rank;
    cudaError_t e = cudaGetDevice(&m->device);
    const size_t bytes = (size_t)world * kMailboxRing * kMailboxSlotWords * 8;
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->local), bytes);
    if (e == cudaSuccess) e = cudaMemset(m->local, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_timeout), 4);
    if (e == cudaSuccess) e = cudaMemset(m->d_timeout, 0, 4);
    if (e != cudaSuccess) {
        g_last_cuda = e;
        (void)cudaGetLastError();
        hd_mailbox_destroy(m);
        return HD_ERR_CUDA;
    }
    m->peers.buf[rank] = m->local;
    *out = m;
    return HD_OK;
}

hd_status hd_mailbox_export(hd_mailbox *m, uint8_t handle_out[64]) {
    if (!m || !handle_out) return HD_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    HD_CUDA(cudaIpcGetMemHandle(&h, m->local));
    std::memcpy(handle_out, &h, 64);
    return HD_OK;
}

hd_status hd_mailbox_connect(hd_mailbox *m, const uint8_t *handles) {
    if (!m || !handles) return HD_ERR_INVALID;
    for (uint32_t p = 0; p < m->world; ++p) {
        if (p == m->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)p * 64, 64);
        void *ptr = nullptr;
        HD_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        m->peers.buf[p] = static_cast<unsigned long long *>(ptr);
        m->opened[p] = true;
    }
    return HD_OK;
}

hd_status hd_mailbox_publish(hd_mailbox *m, const uint64_t *d_info3, uint64_t seq, void *stream) {
    if (!m || !d_info3 || seq == 0) return HD_ERR_INVALID;
    hd_mailbox_publish_kernel<<<1, kMailboxMaxWorld, 0, static_cast<cudaStream_t>(stream)>>>(
        m->peers, m->world, m->rank, reinterpret_cast<const unsigned long long *>(d_info3), seq);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

hd_status hd_mailbox_wait(hd_mailbox *m, uint64_t seq, uint64_t *d_all_info3, void *stream) {
    if (!m || !d_all_info3 || seq == 0) return HD_ERR_INVALID;
    hd_mailbox_wait_kernel<<<1, kMailboxMaxWorld, 0, static_cast<cudaStream_t>(stream)>>>(
        m->local, m->world, seq, reinterpret_cast<unsigned long long *>(d_all_info3), m->d_timeout);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// 1 if a wait gave up because a peer never published (checked by the caller after it synchronised the stream)
int hd_mailbox_timed_out(hd_mailbox *m) {
    if (!m) return 1;
    uint32_t v = 0;
    if (cudaMemcpy(&v, m->d_timeout, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return v != 0;
}

// ------------------------------------------------------------------ tuning
void hd_set_trace(void *) {}   // (the per-step trace of the round-1 kernel is gone; kept for ABI stability)
void hd_set_tuning(int segs_per_thread, int threads_per_tile) {
    // warp-tile kernel: threads_per_tile = decode threads per CTA (a multiple of 32, at most 27 warps)
    // segs_per_thread < 0: also keep fixed-length codes on the general kernels (tests / A-B timing of hd_fixed.cuh)
    g_fixed_fast_path = segs_per_thread >= 0;
    g_tile_segs = segs_per_thread > 0 ? segs_per_thread : 0;
    g_tile_warps = (threads_per_tile > 0 && threads_per_tile % 32 == 0 && threads_per_tile <= (int)kTileMaxWarps * 32) ? threads_per_tile / 32 : 0;
}

uint32_t hd_table_fixed_length(const hd_table *table) {
    uint32_t L = 0;
    return (table && g_fixed_fast_path && fixed_length_code(table, L)) ? L : 0u;
}

// ---------------------------------------------------------------- profiling
uint64_t hd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void hd_profile_enable(int on) {
    if (on && !g_prof.created) {
        for (int i = 0; i < HD_PROFILE_SLOTS; ++i)
            for (int j = 0; j < 2; ++j) cudaEventCreate(&g_prof.ev[i][j]);
        g_prof.created = true;
    }
    g_prof.on = on != 0;
    for (int i = 0; i < HD_PROFILE_SLOTS; ++i) g_prof.used[i] = false;
}

hd_status hd_profile_read(float ms_out[HD_PROFILE_SLOTS]) {
    if (!ms_out || !g_prof.created) return HD_ERR_INVALID;
    for (int i = 0; i < HD_PROFILE_SLOTS; ++i) {
        ms_out[i] = 0.f;
        if (!g_prof.used[i]) continue;
        HD_CUDA(cudaEventSynchronize(g_prof.ev[i][1]));
        HD_CUDA(cudaEventElapsedTime(&ms_out[i], g_prof.ev[i][0], g_prof.ev[i][1]));
        g_prof.used[i] = false;
    }
    return HD_OK;
}

// ------------------------------------------------------- host-buffer decoder
// Persistent device buffers + one stream: the per-call work is H2D, decode,
// D2H -- the region gpuhd/src/demo.cc:170-197 ("Decompression time") and
// decoder.cu:760-803 time, minus their per-call cudaMalloc.
constexpr int kHostPieces = 24;          // upper bound of pipeline pieces per host-buffer decode
constexpr uint64_t kHostPieceSegs = 1 << 20;   // 16 MiB of stream per full piece (a multiple of every tile size); measured: 8 MiB pieces are slower (per-piece launch overhead), PCIe in both directions at once is the floor
constexpr uint64_t kHostHaloSegs = 4096;  // left halo of a self-sync piece (64 KiB of stream)

struct hd_host_decoder {
    int device = 0;
    uint64_t max_units = 0, max_symbols = 0;
    uint8_t *d_units = nullptr, *d_gaps = nullptr, *d_out = nullptr, *d_ws = nullptr;
    size_t ws_bytes = 0;
    cudaStream_t stream = nullptr;                 // compute
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // copy engines: uploads / downloads overlap the decode
    cudaEvent_t ev_piece[kHostPieces] = {};         // "piece i is on the device"
    cudaEvent_t ev_dec[kHostPieces] = {};           // "piece i is decoded and its info is in h_info"
    uint64_t *d_info = nullptr;                     // per piece {count, entry, exit, output offset after the piece}
    uint64_t *h_info = nullptr;                     // pinned copy
    uint32_t *h_flags = nullptr;                    // pinned, per piece: self-synchronisation did not converge blind
    hd_table *cached_table = nullptr;               // device tables of the last code seen (rebuilt only when it changes)
    std::vector<uint8_t> cached_key;
};

void hd_host_decoder_destroy(hd_host_decoder *h) {
    if (!h) return;
    if (h->d_units) cudaFree(h->d_units);
    if (h->d_gaps) cudaFree(h->d_gaps);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_ws) cudaFree(h->d_ws);
    if (h->d_info) cudaFree(h->d_info);
    if (h->h_info) cudaFreeHost(h->h_info);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    hd_table_free(h->cached_table);
    for (auto &e : h->ev_piece) if (e) cudaEventDestroy(e);
    for (auto &e : h->ev_dec) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    delete h;
}

hd_status hd_host_decoder_create(int device, uint64_t max_units, uint64_t max_symbols, hd_host_decoder **out) {
    if (!out || max_units == 0) return HD_ERR_INVALID;
    if (device >= 0) HD_CUDA(cudaSetDevice(device));
    hd_host_decoder *h = new (std::nothrow) hd_host_decoder();
    if (!h) return HD_ERR_INVALID;
    cudaGetDevice(&h->device);
    h->max_units = max_units;
    h->max_symbols = max_symbols;
    h->ws_bytes = ws_bytes(max_units, true);
    const uint64_t n_seg = div_up(max_units, kSegmentUnits);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&h->d_units), align_up(max_units * 4 + 16, 256));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&h->d_gaps), align_up(div_up(n_seg, 8) * 4 + 16, 256));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&h->d_out), align_up(max_symbols + 16, 256));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&h->d_ws), h->ws_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&h->d_info), kHostPieces * 32);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&h->h_info), kHostPieces * 32, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&h->h_flags), kHostPieces * 4, cudaHostAllocDefault);
    for (int i = 0; i < kHostPieces && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_piece[i], cudaEventDisableTiming);
    for (int i = 0; i < kHostPieces && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_dec[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        g_last_cuda = e;
        (void)cudaGetLastError();
        hd_host_decoder_destroy(h);
        return HD_ERR_CUDA;
    }
    *out = h;
    return HD_OK;
}

// Host-buffer decode as a three-stage pipeline over pieces of the stream: all uploads are queued on the H2D stream up
// front; piece i is decoded on the compute stream as soon as it has arrived (a self-sync piece is a shard with a 64 KiB
// left halo that is already on the device; a gap-array piece takes its entry offset from the gap array); its output is
// downloaded on the D2H stream while the next pieces decode.  PCIe runs in both directions at once, so the wall time
// approaches max(H2D, D2H) instead of H2D + decode + D2H (measured floor on this box for 100 MiB at redundancy 0.5:
// 2.09 ms with both directions saturated, tools/pcie_probe.py).  Three things keep the downloads -- the longer
// direction -- busy from early on and without gaps:
//   * the first pieces are small (1/8, 1/4, 1/2 of a full 16 MiB piece), so the first download starts early;
//   * the output position of piece i+1 is chained ON THE DEVICE (K4 reads the running offset the previous launch
//     left behind), so pieces are enqueued two ahead and the compute stream never waits for the host;
//   * the host only waits for "piece i decoded" events to learn {count, entry, exit} and queue the download.
// A self-sync piece whose halo did not synchronise it, or whose blind synchronisation did not converge (codes that do
// not self-synchronise), sends the rest of the stream down the serial path: one piece at a time, exact entry offsets.
// timing_ms = {time until the first piece is on the device, from then until the last piece is decoded (uploads and
// downloads overlap it), time of the remaining downloads}.
struct HostPlan {
    int n_pieces;
    uint64_t seg[kHostPieces + 1];   // piece i = segments [seg[i], seg[i+1])
};

static HostPlan host_plan(uint64_t n_seg) {
    HostPlan P;
    static const uint64_t tuned = [] {   // developer knob: HD_HOST_PIECE_LOG2 = log2(segments per full piece)
        const char *e = std::getenv("HD_HOST_PIECE_LOG2");
        const int v = e ? std::atoi(e) : 0;
        return v >= 13 && v <= 24 ? (uint64_t)1 << v : kHostPieceSegs;
    }();
    uint64_t base = tuned;
    if (n_seg > base * (kHostPieces - 3)) base = align_up(div_up(n_seg, kHostPieces - 3), 8192);
    // ramp up (1/8, 1/4, 1/2 of a full piece) so that the first download starts early, then full pieces.  A ramp down
    // at the end (short last download) was measured and is no faster: with the uploads queued up front the two PCIe
    // directions are both busy until the end either way, and more pieces cost more (tools/pcie_probe.py: dependent
    // piecewise copies alone, without any decode, take 2.64 ms for this workload; the whole call takes 2.63 ms).
    const uint64_t ramp[3] = {base / 8, base / 4, base / 2};
    int n = 0;
    uint64_t s = 0;
    P.seg[0] = 0;
    while (s < n_seg) {
        uint64_t take = n < 3 ? ramp[n] : base;
        if (n_seg - s < take + ramp[0]) take = n_seg - s;   // no sliver at the end
        s += take;
        P.seg[++n] = s;
    }
    P.n_pieces = n;
    return P;
}

// serial tail of the pipeline (and the whole of it for codes that do not self-synchronise): pieces i0.. one at a time
// (base = segments of left halo in front of the piece plan: a shard of a longer stream, see host_decode_common)
static hd_status host_decode_serial(hd_host_decoder *h, const HostPlan &P, int i0, uint64_t offset, uint64_t prev_exit,
                                    uint64_t n_units, const uint8_t *gap_words, const hd_table *tab, uint8_t *out,
                                    uint64_t n_symbols, uint64_t base = 0, uint32_t first_bit0 = 0,
                                    uint64_t *info_out = nullptr) {
    uint32_t *d_units = reinterpret_cast<uint32_t *>(h->d_units);
    uint32_t *d_gaps = reinterpret_cast<uint32_t *>(h->d_gaps);
    for (int i = i0; i < P.n_pieces && offset < n_symbols; ++i) {
        const uint64_t s0 = P.seg[i], s1 = P.seg[i + 1];
        const uint64_t b0 = base + s0, b1 = base + s1;   // segments in the buffer
        const uint64_t shard_units = std::min<uint64_t>(n_units, b1 * 4) - b0 * 4;
        const uint64_t readable_end = std::min<uint64_t>(n_units, b1 * 4 + 4);
        HD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_piece[i], 0));
        hd_status st;
        if (gap_words) {
            uint32_t first_bit = first_bit0;
            if (s0) {
                const uint64_t j = s0 - 1;
                uint32_t w;
                std::memcpy(&w, gap_words + (j >> 3) * 4, 4);
                first_bit = (w >> (4u * (uint32_t)(j & 7))) & 15u;
            }
            st = hd_decode_gap_ex(d_units + b0 * 4, readable_end - b0 * 4, d_gaps + s0 / 8, s1 - s0, first_bit, tab,
                                  h->d_out + offset, n_symbols - offset, h->d_ws, h->ws_bytes, h->d_info, h->stream);
        } else {
            const uint64_t halo = s0 ? std::min<uint64_t>(kHostHaloSegs, b0) / 8 * 8 : base;
            st = hd_decode_selfsync_shard(d_units + (b0 - halo) * 4, halo * 4, shard_units,
                                          readable_end - (b0 - halo) * 4, 0, tab, h->d_out + offset, n_symbols - offset,
                                          h->d_ws, h->ws_bytes, h->d_info, h->stream);
        }
        if (st != HD_OK) return st;
        HD_CUDA(cudaMemcpyAsync(h->h_info, h->d_info, 24, cudaMemcpyDeviceToHost, h->stream));
        HD_CUDA(cudaStreamSynchronize(h->stream));
        if (!gap_words && i > 0 && h->h_info[1] != prev_exit) {
            // the halo did not synchronise this piece (a code that does not self-synchronise within 64 KiB): decode
            // it again from the exact offset at which the previous piece's parse left
            st = hd_decode_selfsync_shard(d_units + b0 * 4, 0, shard_units, readable_end - b0 * 4, (uint32_t)(prev_exit & 15u), tab,
                                          h->d_out + offset, n_symbols - offset, h->d_ws, h->ws_bytes, h->d_info,
                                          h->stream);
            if (st != HD_OK) return st;
            HD_CUDA(cudaMemcpyAsync(h->h_info, h->d_info, 24, cudaMemcpyDeviceToHost, h->stream));
            HD_CUDA(cudaStreamSynchronize(h->stream));
        }
        const uint64_t count = h->h_info[0];
        if (info_out && i == 0) info_out[1] = h->h_info[1] & 0xffffu;
        prev_exit = h->h_info[2];
        HD_CUDA(cudaMemcpyAsync(out + offset, h->d_out + offset, count, cudaMemcpyDeviceToHost, h->s_d2h));
        offset += count;
    }
    HD_CUDA(cudaStreamSynchronize(h->s_d2h));
    HD_CUDA(cudaStreamSynchronize(h->s_h2d));
    if (info_out) {   // a shard: the caller learns the count
        info_out[0] = offset;
        info_out[2] = prev_exit;
        return HD_OK;
    }
    return offset < n_symbols ? HD_ERR_FORMAT : HD_OK;   // FORMAT: the stream holds fewer codewords than asked for
}

// A SHARD of a longer stream (multi-GPU, hd_host_decode_*_shard) is the same pipeline with three differences: the
// buffer starts with `base` segments of left halo that are uploaded with the first piece and only parsed (self-sync) --
// the first piece's entry offset is then speculative and reported, not checked; only n_dec_segs segments are decoded;
// and the number of symbols is an output (info_out = {count, entry offset of the first piece, exit offset}).
static hd_status host_decode_common(hd_host_decoder *h, const uint32_t *units, uint64_t n_units,
                                    const uint8_t *gap_words, uint64_t n_segments, const hd_table *tab,
                                    uint8_t *out, uint64_t n_symbols, float timing_ms[3], uint64_t base = 0,
                                    uint64_t n_dec_segs = 0, uint32_t first_bit0 = 0, uint64_t *info_out = nullptr) {
    if (n_units > h->max_units || n_symbols > h->max_symbols) return HD_ERR_INVALID;
    HD_CUDA(cudaSetDevice(h->device));
    uint32_t *d_units = reinterpret_cast<uint32_t *>(h->d_units);
    uint32_t *d_gaps = reinterpret_cast<uint32_t *>(h->d_gaps);
    const uint64_t n_seg = n_dec_segs ? n_dec_segs : div_up(n_units, kSegmentUnits);
    const HostPlan P = host_plan(n_seg);
    const auto t0 = std::chrono::steady_clock::now();

    // ---- stage 1: every upload, one event per piece (a piece also brings the next piece's first unit along: the
    // last codeword of a piece may end there; re-uploading those 16 bytes is harmless)
    if (gap_words) {
        HD_CUDA(cudaMemcpyAsync(d_gaps, gap_words, div_up(n_segments, 8) * 4, cudaMemcpyHostToDevice, h->s_h2d));
    }
    for (int i = 0; i < P.n_pieces; ++i) {
        const uint64_t u0 = i ? (base + P.seg[i]) * 4 : 0;   // (the first piece brings the shard's left halo along)
        const uint64_t u1 = std::min<uint64_t>(n_units, (base + P.seg[i + 1]) * 4 + 4);
        HD_CUDA(cudaMemcpyAsync(d_units + u0, units + u0, (u1 - u0) * 4, cudaMemcpyHostToDevice, h->s_h2d));
        HD_CUDA(cudaEventRecord(h->ev_piece[i], h->s_h2d));
    }

    // ---- stage 2: decode piece i into d_out at the offset the previous piece left on the device
    auto enqueue_piece = [&](int i) -> hd_status {
        const uint64_t s0 = P.seg[i], s1 = P.seg[i + 1];
        const uint64_t b0 = base + s0, b1 = base + s1;   // segments in the buffer
        const uint64_t shard_units = std::min<uint64_t>(n_units, b1 * 4) - b0 * 4;
        const uint64_t readable_end = std::min<uint64_t>(n_units, b1 * 4 + 4);
        uint64_t *info = h->d_info + 4 * i;
        const uint64_t *off_in = i ? h->d_info + 4 * (i - 1) + 3 : nullptr;
        HD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_piece[i], 0));
        hd_status st;
        if (gap_words) {
            uint32_t first_bit = first_bit0;
            if (s0) {
                const uint64_t j = s0 - 1;
                uint32_t w;
                std::memcpy(&w, gap_words + (j >> 3) * 4, 4);
                first_bit = (w >> (4u * (uint32_t)(j & 7))) & 15u;
            }
            const uint64_t piece_units = readable_end - b0 * 4;
            Workspace w = carve(h->d_ws, piece_units, false);
            // (info doubles as the count word and as the {count, entry, exit} triple: a shard reports its exit offset)
            st = run_gap_decode(d_units + b0 * 4, piece_units, d_gaps + s0 / 8, s1 - s0, first_bit, tab, h->d_out,
                                n_symbols, w, info, h->stream, 0, nullptr, info, false, off_in, info + 3);
            h->h_flags[i] = 0;
        } else {
            const uint64_t halo = s0 ? std::min<uint64_t>(kHostHaloSegs, b0) / 8 * 8 : base;
            st = shard_enqueue_impl(d_units + (b0 - halo) * 4, halo * 4, shard_units, readable_end - (b0 - halo) * 4, 0,
                                    tab, h->d_out, n_symbols, h->d_ws, h->ws_bytes, info, h->h_flags + i, h->stream,
                                    off_in, info + 3);
        }
        if (st != HD_OK) return st;
        HD_CUDA(cudaMemcpyAsync(h->h_info + 4 * i, info, 32, cudaMemcpyDeviceToHost, h->stream));
        HD_CUDA(cudaEventRecord(h->ev_dec[i], h->stream));
        return HD_OK;
    };
    int enq = 0;
    for (; enq < P.n_pieces && enq < 2; ++enq) {
        hd_status st = enqueue_piece(enq);
        if (st != HD_OK) return st;
    }
    HD_CUDA(cudaEventSynchronize(h->ev_piece[0]));
    const auto t1 = std::chrono::steady_clock::now();

    // ---- stage 3: as pieces finish, queue their downloads and keep two pieces enqueued ahead
    uint64_t offset = 0, prev_exit = 0;
    for (int i = 0; i < P.n_pieces; ++i) {
        HD_CUDA(cudaEventSynchronize(h->ev_dec[i]));
        const uint64_t *info = h->h_info + 4 * i;
        if (!gap_words && (h->h_flags[i] != 0 || (i > 0 && (info[1] & 0xffffu) != (prev_exit & 0xffffu)))) {
            HD_CUDA(cudaStreamSynchronize(h->stream));   // pieces enqueued ahead decoded to a wrong position: drop them
            const hd_status st = host_decode_serial(h, P, i, offset, prev_exit, n_units, gap_words, tab, out, n_symbols,
                                                    base, first_bit0, info_out);
            if (timing_ms) {
                timing_ms[0] = (float)std::chrono::duration<double, std::milli>(t1 - t0).count();
                timing_ms[1] = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
                timing_ms[2] = 0.f;
            }
            return st;
        }
        const uint64_t count = info[0];
        if (info_out && i == 0) info_out[1] = info[1] & 0xffffu;
        prev_exit = info[2];
        if (count) HD_CUDA(cudaMemcpyAsync(out + offset, h->d_out + offset, count, cudaMemcpyDeviceToHost, h->s_d2h));
        offset += count;
        if (enq < P.n_pieces) {
            hd_status st = enqueue_piece(enq++);
            if (st != HD_OK) return st;
        }
    }
    const auto t2 = std::chrono::steady_clock::now();
    HD_CUDA(cudaStreamSynchronize(h->s_d2h));
    HD_CUDA(cudaStreamSynchronize(h->s_h2d));
    if (info_out) {   // a shard: the caller learns the count
        info_out[0] = offset;
        info_out[2] = prev_exit;
    } else if (offset < n_symbols) {
        return HD_ERR_FORMAT;   // the stream holds fewer codewords than asked for
    }
    if (timing_ms) {
        timing_ms[0] = (float)std::chrono::duration<double, std::milli>(t1 - t0).count();
        timing_ms[1] = (float)std::chrono::duration<double, std::milli>(t2 - t1).count();
        timing_ms[2] = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count();
    }
    return HD_OK;
}

hd_status hd_host_decode_selfsync(hd_host_decoder *h, const uint32_t *units, uint64_t n_units,
                                  const hd_lut_entry *lut, uint32_t lut_bits, uint8_t *out, uint64_t n_symbols,
                                  float timing_ms[3]) {
    if (!h || !units || !lut || (!out && n_symbols)) return HD_ERR_INVALID;
    if (n_units == 0) return n_symbols == 0 ? HD_OK : HD_ERR_INVALID;
    HD_CUDA(cudaSetDevice(h->device));
    const size_t key_bytes = sizeof(hd_lut_entry) << lut_bits;
    if (!h->cached_table || h->cached_key.size() != key_bytes + 1 || h->cached_key[0] != 0 ||
        std::memcmp(h->cached_key.data() + 1, lut, key_bytes) != 0) {
        hd_table_free(h->cached_table);
        h->cached_table = nullptr;
        hd_status st = hd_table_from_flat_lut(lut, lut_bits, h->stream, &h->cached_table);
        if (st != HD_OK) return st;
        HD_CUDA(cudaStreamSynchronize(h->stream));
        h->cached_key.assign(key_bytes + 1, 0);
        std::memcpy(h->cached_key.data() + 1, lut, key_bytes);
    }
    return host_decode_common(h, units, n_units, nullptr, 0, h->cached_table, out, n_symbols, timing_ms);
}

hd_status hd_host_decode_huff(hd_host_decoder *h, const void *file, size_t size, uint8_t *out, size_t out_cap,
                              size_t *out_size, float timing_ms[3]) {
    if (!h) return HD_ERR_INVALID;
    hd_huff_view v;
    hd_status st = hd_huff_parse(file, size, &v);
    if (st != HD_OK) return st;
    if (out_size) *out_size = v.original_size;
    if (!out || out_cap < v.original_size) return HD_ERR_INVALID;
    if (v.original_size == 0 || v.n_units == 0) return HD_OK;
    HD_CUDA(cudaSetDevice(h->device));
    const size_t key_bytes = (size_t)v.symbol_count * sizeof(hd_sym_len);
    if (!h->cached_table || h->cached_key.size() != key_bytes + 1 || h->cached_key[0] != 1 ||
        std::memcmp(h->cached_key.data() + 1, v.symbols, key_bytes) != 0) {
        hd_table_free(h->cached_table);
        h->cached_table = nullptr;
        st = hd_table_from_lengths(v.symbols, (uint32_t)v.symbol_count, h->stream, &h->cached_table);
        if (st != HD_OK) return st;
        HD_CUDA(cudaStreamSynchronize(h->stream));
        h->cached_key.assign(key_bytes + 1, 1);
        std::memcpy(h->cached_key.data() + 1, v.symbols, key_bytes);
    }
    // the payload may be only 2-byte aligned in the file image; cudaMemcpy does not care
    return host_decode_common(h, reinterpret_cast<const uint32_t *>(v.units), v.n_units, v.gap_words, v.n_segments,
                              h->cached_table, out, v.original_size, timing_ms);
}

// One rank's shard of a longer stream from / to HOST buffers (multi-GPU end to end: every rank runs the same
// three-stage pipeline over its shard; hd_mgpu_host_step adds the exchange of the counts).  units -> [left halo | shard
// | right halo] as hd_mgpu_plan_shard lays it out; info_out = {symbols written to out, entry offset, exit offset}.
hd_status hd_host_decode_selfsync_shard(hd_host_decoder *h, const uint32_t *units, uint64_t n_units_halo,
                                        uint64_t n_units_shard, uint64_t n_units_readable, const hd_table *table,
                                        uint8_t *out, uint64_t out_cap, uint64_t info_out[3], float timing_ms[3]) {
    if (!h || !units || !table || !out || !info_out || n_units_shard == 0) return HD_ERR_INVALID;
    if (n_units_halo % (kSegmentUnits * 8) || n_units_readable < n_units_halo + n_units_shard) return HD_ERR_INVALID;
    return host_decode_common(h, units, n_units_readable, nullptr, 0, table, out, std::min<uint64_t>(out_cap, h->max_symbols),
                              timing_ms, n_units_halo / kSegmentUnits, div_up(n_units_shard, kSegmentUnits), 0, info_out);
}

// gap_words: the shard's own elements (element 0 = entry offset of the shard's segment 1); entry_bit = offset at
// which the shard's first segment starts (decoder.cu:505-507: the gap element before the shard)
hd_status hd_host_decode_gap_shard(hd_host_decoder *h, const uint32_t *units, uint64_t n_units_readable,
                                   const uint32_t *gap_words, uint64_t n_segments, uint32_t entry_bit,
                                   const hd_table *table, uint8_t *out, uint64_t out_cap, uint64_t info_out[3],
                                   float timing_ms[3]) {
    if (!h || !units || !gap_words || !table || !out || !info_out || n_segments == 0 || entry_bit > 15) return HD_ERR_INVALID;
    if (n_segments > div_up(n_units_readable, kSegmentUnits)) return HD_ERR_INVALID;
    return host_decode_common(h, units, n_units_readable, reinterpret_cast<const uint8_t *>(gap_words), n_segments, table, out,
                              std::min<uint64_t>(out_cap, h->max_symbols), timing_ms, 0, n_segments, entry_bit, info_out);
}

// device buffers of a host decoder (after a call they hold the uploaded units and the decoded output): lets the
// multi-GPU layer repeat a shard decode on the device in the rare repair case
void hd_host_decoder_buffers(hd_host_decoder *h, uint32_t **d_units, uint32_t **d_gap_words, uint8_t **d_out,
                             void **d_workspace, size_t *workspace_bytes) {
    if (!h) return;
    if (d_units) *d_units = reinterpret_cast<uint32_t *>(h->d_units);
    if (d_gap_words) *d_gap_words = reinterpret_cast<uint32_t *>(h->d_gaps);
    if (d_out) *d_out = h->d_out;
    if (d_workspace) *d_workspace = h->d_ws;
    if (workspace_bytes) *workspace_bytes = h->ws_bytes;
}

// One-shot `./bin/demo` shape (gpuhd/src/demo.cc:155-197): allocate, H2D, decode, D2H, free.
hd_status hd_selfsync_decode_host(const uint32_t *units, uint64_t n_units, const hd_lut_entry *lut,
                                  uint32_t lut_bits, uint8_t *out, uint64_t n_symbols, int device,
                                  float timing_ms[3]) {
    hd_host_decoder *h = nullptr;
    hd_status st = hd_host_decoder_create(device, n_units ? n_units : 1, n_symbols, &h);
    if (st != HD_OK) return st;
    st = hd_host_decode_selfsync(h, units, n_units, lut, lut_bits, out, n_symbols, timing_ms);
    hd_host_decoder_destroy(h);
    return st;
}

// Whole-file convenience with host buffers: the `./bin/decoder in out` shape of
// Huffman_coding_Gap_arrays/decoder/src/huff.cpp:31-129 + decoder.cu:732-815
// (allocate, H2D, one decode launch instead of 200, D2H).
hd_status hd_huff_decode_host(const void *file, size_t size, uint8_t *out, size_t out_cap, size_t *out_size,
                              int device, float timing_ms[3]) {
    hd_huff_view v;
    hd_status st = hd_huff_parse(file, size, &v);
    if (st != HD_OK) return st;
    if (out_size) *out_size = v.original_size;
    if (!out || out_cap < v.original_size) return HD_ERR_INVALID;
    if (device >= 0) HD_CUDA(cudaSetDevice(device));
    if (v.original_size == 0 || v.n_units == 0) return HD_OK;
    hd_table *tab = nullptr;
    st = hd_table_from_lengths(v.symbols, (uint32_t)v.symbol_count, nullptr, &tab);
    if (st != HD_OK) return st;
    const size_t gap_bytes = align_up(((size_t)v.n_segments + 7) / 8 * 4, 256);
    const size_t unit_bytes = align_up((size_t)v.n_units * 4, 256);
    const size_t ws = ws_bytes(v.n_units, false);
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_ws = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    auto cleanup = [&]() {
        if (d_in) cudaFree(d_in);
        if (d_out) cudaFree(d_out);
        if (d_ws) cudaFree(d_ws);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        hd_table_free(tab);
    };
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d_in), gap_bytes + unit_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&d_out), align_up(v.original_size, 256));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&d_ws), ws);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ev[i]);
    if (e != cudaSuccess) { g_last_cuda = e; (void)cudaGetLastError(); cleanup(); return HD_ERR_CUDA; }
    cudaEventRecord(ev[0], nullptr);
    // gap words and units are uploaded to separately aligned regions (the file
    // payload is only 2-byte aligned when the symbol count is odd)
    e = cudaMemcpyAsync(d_in, v.gap_words, ((size_t)v.n_segments + 7) / 8 * 4, cudaMemcpyHostToDevice, nullptr);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in + gap_bytes, v.units, (size_t)v.n_units * 4, cudaMemcpyHostToDevice, nullptr);
    cudaEventRecord(ev[1], nullptr);
    if (e == cudaSuccess) {
        st = hd_decode_gap(reinterpret_cast<uint32_t *>(d_in + gap_bytes), v.n_units, reinterpret_cast<uint32_t *>(d_in),
                           v.n_segments, tab, d_out, v.original_size, d_ws, ws, nullptr);
        if (st != HD_OK) { cleanup(); return st; }
    }
    cudaEventRecord(ev[2], nullptr);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, v.original_size, cudaMemcpyDeviceToHost, nullptr);
    cudaEventRecord(ev[3], nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) { g_last_cuda = e; (void)cudaGetLastError(); cleanup(); return HD_ERR_CUDA; }
    if (timing_ms) for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&timing_ms[i], ev[i], ev[i + 1]);
    cleanup();
    return HD_OK;
}

}  // extern "C"
