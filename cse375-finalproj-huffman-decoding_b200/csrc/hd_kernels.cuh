// hd_kernels.cuh -- device code of the B200 (sm_100a) Huffman decoder.
//
//   K4  hd_decode_write_kernel : gap-array segment decode + single-pass
//        decoupled look-back prescan + decode-and-write, one persistent launch.
//        Replaces gpu_dec_l1_l2 (Huffman_coding_Gap_arrays/decoder/src/
//        decoder.cu:454-730) and phases 3+4 of CUHD (gpuhd/src/
//        cuhd_gpu_decoder.cu:329-420 + thrust scan :498-505).
//   K1a hd_sync_parse_kernel   : speculative subsequence synchronisation: every
//        thread parses a halo + its chunk, seams are repaired inside the warp
//        (replaces phase 1, cuhd_gpu_decoder.cu:145-233)
//   K1b hd_sync_walk_kernel    : seams between warp tiles, walked to the merge
//        (replaces phase 2 + host loop, cuhd_gpu_decoder.cu:235-327, 461-495)
//
// Integer/byte work only: no tensor cores.  See DESIGN.md for the data layout
// and the per-kernel roofline.
#ifndef HD_KERNELS_CUH_
#define HD_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "hd_core.h"
#include "hd_internal.h"

namespace hd {

// ------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint64_t ld_relaxed_gpu_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_cta_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
// order-independent publication of a tile status: a PREFIX word (flag 2) always beats an AGGREGATE word (flag 1)
__device__ __forceinline__ void red_max_gpu_u64(uint64_t *p, uint64_t v) {
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// TMA bulk store shared::cta -> global (SASS: UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async (TMA) proxy
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory transaction barrier used as a one-producer flag: waiters are suspended by the hardware instead of
// spinning on a load (a spinning warp takes issue slots from the decode warps of its scheduler)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!done);
}

// ------------------------------------------------------------ kernel params
constexpr uint64_t kFlagAggregate = 1ull << 62;   // tile aggregate published
constexpr uint64_t kFlagPrefix = 2ull << 62;      // inclusive prefix published
constexpr uint64_t kFlagMask = 3ull << 62;
// One status word per 32-byte sector: every scan warp reads the statuses of the ~#SMs tiles in flight below its own,
// so densely packed words turn a few L2 lines into a hot spot (measured: ~2.5 us per 256-word poll).
constexpr uint32_t kStatusStride = 4;   // in 64-bit words

struct TableDev {
    const uint8_t *blob;      // device copy of TableBlob::bytes
    uint32_t lut_bits;
    uint32_t has_long;
    uint32_t off_sync, off_len, off_order, off_long;
};

struct DecodeParams {
    const uint32_t *units;
    uint64_t n_units;            // units that may be READ (includes a right halo when the stream continues)
    const uint32_t *gap_words;   // element i = start offset of segment i+1
    uint64_t n_segments;         // segments to DECODE
    uint32_t first_bit;          // start offset of segment 0 ...
    const uint32_t *entry_word;  // ... unless set: start offset of segment 0 = *entry_word >> 28
    uint64_t *shard_info;        // optional {count, entry_bit, exit_bit} for the multi-GPU seam check
    TableDev table;
    uint8_t *out;
    uint64_t n_symbols;
    const uint64_t *out_offset_in;   // optional: decode to out + *out_offset_in (output capacity n_symbols counts from `out`);
    uint64_t *out_offset_out;        // optional: receives *out_offset_in + symbols written -- chains the pieces of a
                                     // pipelined host-buffer decode on the device, without a host round trip per piece
    uint64_t *tile_status;       // n_tiles words, zeroed before launch
    uint32_t *tile_counter;      // zeroed before launch
    uint64_t n_tiles;
    uint64_t *count_out;         // optional: total symbols written
    uint32_t stage_words;        // staging rows (one 32-bit word per thread each)
    uint32_t tile_bytes;         // output tile bytes (multiple of 16)
    uint32_t tile_buffers;       // 2: tiles of consecutive steps alternate (no decode-warp barrier); 1: one barrier per step
    unsigned long long *trace;   // optional (hd_set_trace): per-step timestamps of CTA kTraceCta, 8 u64 per step
};
constexpr uint32_t kTraceCta = 5, kTraceSteps = 48;
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct SyncParams {
    const uint32_t *units;
    uint64_t n_units;            // units that may be READ
    uint64_t n_segments;
    uint64_t n_chunks;           // ceil(n_segments / kSyncChunkSegs)
    uint64_t n_tiles;            // warp tiles: ceil(n_chunks / 32)
    uint32_t first_bit;
    TableDev table;
    uint32_t *gap_words;         // out: two words per chunk (nibble k of word w = landing into segment 8w+k+1)
    uint8_t *cur_entry;          // per chunk: the entry offset its recorded landings were parsed from
    uint32_t *changed;           // per-round "some chunk's exit changed" counter
    uint32_t *tile_counter;      // ticket counter of the persistent K1a warps (zeroed before launch)
    uint32_t *tile_exit;         // per K1a warp tile: kExitPublished | offset at which its parse leaves it (zeroed before launch)
};
constexpr uint32_t kExitPublished = 0x100u;
constexpr uint32_t kExitWaitPolls = 1u << 16;   // bounded wait for a predecessor tile (then: flag `changed`, K1b rounds decide)

// =========================================================================
// K4: decode-and-write.  ONE persistent CTA per SM: NT decode threads (one per
// CHUNK of SEGS consecutive 128-bit segments; NT chunks = one tile) plus one
// scan warp.  Tile ids come from an atomic ticket counter so that the look-back
// never waits on a tile that is not running.  The CTA is software pipelined over
// two staging buffers so that the L2 round trips of the decoupled look-back are
// hidden behind the decode of the next tile:
//
//   decode warps, step k:   load + decode tile k -> staging[k&1]   (decode_chunk)
//                           warp-scan the symbol counts, arrive FULL[k&1]
//                           wait READY[(k-1)&1]  (prefix of tile k-1, long done)
//                           compact staging[(k-1)&1] -> output tile, flush it
//                           (one TMA bulk store, SASS UBLKCP, + scalar head/tail)
//   scan warp, step k:      wait FULL[k&1]; take the ticket for step k+2; scan
//                           the warp totals (the tile aggregate was published
//                           by the last decode warp to finish); look back over
//                           the predecessors' aggregates/prefixes (64-bit
//                           flagged words, ld.relaxed.gpu / red.max); publish
//                           the inclusive prefix; arrive READY[k&1]
//
// FULL is a named barrier (decode warps bar.arrive, the scan warp bar.sync),
// READY an mbarrier the decode warps wait on: no decode warp ever waits for
// another decode warp (with two output tiles), only for the scan warp.
// =========================================================================
__device__ __forceinline__ uint32_t gap_nibble(const uint32_t *gap_words, uint64_t j) {
    return (__ldg(gap_words + (j >> 3)) >> (4u * (uint32_t)(j & 7))) & 15u;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t n) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
constexpr uint32_t kBarFull = 1, kBarReady = 3, kBarDecode = 5;

// Decoupled look-back, 256 predecessors per round (8 per lane, nearest first).
// All CTAs of the persistent grid work on neighbouring tiles at the same time, so
// the nearest published PREFIX is typically one "generation" (~#SMs tiles) back:
// a 32-wide window would need ~5 serial L2 round trips per tile, this needs one.
// Only slots nearer than the nearest prefix are waited for.
__device__ __forceinline__ uint64_t lookback_wide(const uint64_t *status, uint64_t tile, uint32_t lane, uint32_t *polls = nullptr) {
    uint64_t excl = 0;
    uint32_t n_polls = 0;
    int64_t hi = (int64_t)tile - 1;
    for (;;) {
        const int64_t first = hi - 8 * (int64_t)lane;
        ++n_polls;
        uint64_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t idx = first - j;
            v[j] = idx >= 0 ? ld_relaxed_gpu_u64(status + idx * kStatusStride) : kFlagPrefix;   // virtual tile -1: prefix 0
        }
        uint64_t sum = 0;
        uint32_t state = 0;   // 0: aggregates only, 1: reached a prefix, 2: reached an unpublished slot first
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (state == 0) {
                const uint64_t f = v[j] & kFlagMask;
                if (f == 0) {
                    state = 2;
                } else {
                    sum += v[j] & ~kFlagMask;
                    if (f == kFlagPrefix) state = 1;
                }
            }
        }
        const uint32_t m_found = __ballot_sync(0xffffffffu, state == 1);
        const uint32_t m_wait = __ballot_sync(0xffffffffu, state == 2);
        const uint32_t stop_lane = m_found ? (uint32_t)__ffs(m_found) - 1u : 32u;
        const uint32_t upto = stop_lane >= 31u ? 0xffffffffu : ((2u << stop_lane) - 1u);   // lanes 0..stop_lane
        if (m_wait & upto) continue;   // something nearer than the nearest prefix is not published yet: poll again
        uint64_t c = (lane <= stop_lane) ? sum : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (m_found) break;
        hi -= 256;
    }
    if (polls) *polls = n_polls;
    return excl;
}

template <int NT, int SEGS, bool LONG>
__global__ void __launch_bounds__(NT + 32, 1) hd_decode_write_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr uint32_t kNW = NT / 32;
    constexpr uint32_t kRowBytes = NT * 4;
    constexpr uint32_t kInRows = 4 * SEGS + 3;
    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    const uint32_t lm = p.table.lut_bits;
    const uint32_t lut_bytes = 4u << lm;
    // where this launch's output starts (chained pieces: written by the previous launch on the stream)
    const uint64_t out_off = p.out_offset_in ? *p.out_offset_in : 0ull;
    uint8_t *const out = p.out + out_off;
    const uint64_t n_symbols = p.n_symbols > out_off ? p.n_symbols - out_off : 0ull;

    uint32_t *s_lut = reinterpret_cast<uint32_t *>(smem);
    uint8_t *s_order = smem + lut_bytes;
    LongTables *s_long = reinterpret_cast<LongTables *>(s_order + 256);
    uint32_t off = lut_bytes + (LONG ? 256u + ((uint32_t)(sizeof(LongTables) + 15) & ~15u) : 0u);
    uint32_t *s_in = reinterpret_cast<uint32_t *>(smem + off);
    off += kInRows * kRowBytes;
    // staging rows must start on a row-size boundary of the shared window (see decode_chunk)
    const uint32_t smem_abs = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t stage_abs = (smem_abs + off + kRowBytes - 1u) & ~(kRowBytes - 1u);
    const uint32_t stage_buf_bytes = p.stage_words * kRowBytes;
    uint8_t *s_stage = smem + (stage_abs - smem_abs);
    uint8_t *s_tile = s_stage + 2u * stage_buf_bytes;   // two output tiles of p.tile_bytes each

    __shared__ uint32_t s_wtot[2][kNW];   // warp totals of the tile in buffer b; the scan warp turns them into exclusive offsets
    __shared__ uint32_t s_tileid[3];   // ring: tile of step k in slot k % 3 (taken by decode thread 0 one step ahead)
    __shared__ unsigned long long s_base[2];
    __shared__ uint32_t s_total[2];
    __shared__ uint32_t s_sum[2], s_arrived[2];   // tile total / number of warps that added theirs
    // s_ready[b] completes its n-th phase once the prefix of step k = b + 2n is in s_base[b].  Parity waits are safe:
    // the scan warp cannot finish step k+2 (the next phase of the same barrier) before every decode warp has arrived
    // at FULL of step k+1, i.e. has long left its wait for step k.
    __shared__ __align__(8) unsigned long long s_ready[2];
    __shared__ uint32_t s_t1_ready;               // the tile of step 1 is in s_tileid[1]

    // Tickets: every tile id comes from the atomic counter, so tiles are only ever claimed by CTAs that are running
    // (no assumption that the whole grid is co-resident).  The FIRST ticket is taken here, before the tables are
    // staged, the SECOND by the scan warp after that (below): by then every CTA has its first tile, so a CTA's first
    // look-back depends only on tiles that are decoded in the same step by the other CTAs, not on second tiles.
    if (t == NT) s_tileid[0] = atomicAdd(p.tile_counter, 1u);
    // ---- stage the tables once per CTA (persistent) ----
    if (t < 2) { s_sum[t] = 0; s_arrived[t] = 0; mbar_init(&s_ready[t], 1u); }
    if (t == 0) s_t1_ready = 0;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.table.blob);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t i = t; i < lut_bytes / 16; i += NT + 32) dst[i] = __ldg(src + i);
        if (LONG) {
            if (t < 64) reinterpret_cast<uint32_t *>(s_order)[t] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_order) + t);
            if (t < sizeof(LongTables) / 4) reinterpret_cast<uint32_t *>(s_long)[t] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_long) + t);
        }
    }
    __syncthreads();

    if (t >= NT) {
        // ======================= scan warp =======================
        if (lane == 0) {
            s_tileid[1] = atomicAdd(p.tile_counter, 1u);
            __threadfence_block();
            st_release_cta_u32(&s_t1_ready, 1u);
        }
        __syncwarp();
        for (uint32_t k = 0;; ++k) {
            const uint32_t b = k & 1u;
            // (slot k % 3 was written by decode thread 0 during step k-1 and is not rewritten before step k+2)
            const uint64_t tile = s_tileid[k % 3u];
            if (tile >= p.n_tiles) break;
            const bool tr = p.trace && blockIdx.x == kTraceCta && k < kTraceSteps && lane == 0;
            named_bar_sync(kBarFull + b, NT + 32);
            // ticket of step k+2 (steps 0 and 1 got theirs in the prologue).  Issued here, but no earlier: a tile must
            // not be claimed long before it is decoded, or every higher tile's look-back waits for it.
            uint32_t ticket = 0;
            if (lane == 0) ticket = atomicAdd(p.tile_counter, 1u);
            if (tr) p.trace[k * 8 + 4] = gtime_ns();
            const uint32_t wt = lane < kNW ? s_wtot[b][lane] : 0u;
            uint32_t wi = wt;
#pragma unroll
            for (int d = 1; d < (int)kNW; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += v;
            }
            if (lane < kNW) s_wtot[b][lane] = wi - wt;   // exclusive warp offsets
            const uint64_t agg = __shfl_sync(0xffffffffu, wi, kNW - 1);
            // (the tile AGGREGATE was already published by the last decode warp to finish; tile 0 published its PREFIX.)
            // Looking back BEFORE the tile is full, while the predecessors are still being decoded, was measured and is
            // slower: 148 scan warps polling the same few L2 lines delay the very publications they wait for.
            uint64_t excl = 0;
            if (HD_WHATIF & 8) excl = tile * 22040ull;   // ~ the true mean of the bench workload (a larger constant would clamp tiles away)
            if (tile > 0 && !(HD_WHATIF & 8)) {
                uint32_t polls = 0;
                excl = lookback_wide(p.tile_status, tile, lane, &polls);
                if (tr) p.trace[k * 8 + 2] = polls;
                if (lane == 0) red_max_gpu_u64(p.tile_status + tile * kStatusStride, kFlagPrefix | (excl + agg));
            }
            if (lane == 0) {
                s_base[b] = excl;
                s_total[b] = (uint32_t)agg;
                if (tile + 1 == p.n_tiles) {
                    uint64_t tot = excl + agg;
                    tot = tot < n_symbols ? tot : n_symbols;
                    if (p.count_out) *p.count_out = tot;
                    if (p.out_offset_out) *p.out_offset_out = out_off + tot;
                    if (p.shard_info) {
                        p.shard_info[0] = tot;
                        p.shard_info[1] = p.entry_word ? (__ldg(p.entry_word) >> 28) : p.first_bit;
                        p.shard_info[2] = gap_nibble(p.gap_words, p.n_segments - 1);   // where the parse leaves the shard
                    }
                }
            }
            if (tr) { p.trace[k * 8 + 5] = gtime_ns(); p.trace[k * 8 + 6] = tile; }
            // slot (k+2)%3 held step k-1's tile, which nobody reads any more
            if (lane == 0) s_tileid[(k + 2u) % 3u] = ticket;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) mbar_arrive(&s_ready[b]);
        }
        return;
    }

    // ======================= decode warps =======================
    DecodeTables T;
    T.lut = s_lut;
    T.idx_shift = 32 - lm;
    T.lut_bits = lm;
    T.one = lm != 0xffffffffu;   // 1 for every real table; ptxas cannot fold it
    T.len_tab = nullptr;
    T.lt = s_long;
    T.order = s_order;
    uint32_t *const col = s_in + t;
    bool store_pending = false, prev_valid = false;
    uint32_t prev_cnt = 0, prev_incl = 0;

    // Inputs of the NEXT tile are prefetched into registers while the previous tile is compacted and flushed,
    // so that the DRAM latency of the unit loads is off the critical path of a step.
    uint4 pf_w[SEGS];
    uint2 pf_la = make_uint2(0u, 0u);
    // raw gap words + nibble shifts; the nibbles are extracted when the step starts so that the prefetch itself
    // never waits for a load.  pf_bits == 0: this thread has no chunk in the tile.
    uint32_t pf_gw0 = 0, pf_gw1 = 0, pf_sh = 0, pf_bits = 0;
    auto prefetch = [&](uint64_t tile) {
        pf_bits = 0;
        if (tile >= p.n_tiles) return;
        const uint64_t g0 = (tile * NT + t) * SEGS;
        if (g0 >= p.n_segments) return;
        const uint64_t left = p.n_segments - g0;
        const uint32_t nseg = left < (uint64_t)SEGS ? (uint32_t)left : (uint32_t)SEGS;
        const uint64_t u0 = g0 * 4;
#pragma unroll
        for (int q = 0; q < SEGS; ++q) {
            const uint64_t u = u0 + 4u * q;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (u + 4 <= p.n_units) {
                w = __ldg(reinterpret_cast<const uint4 *>(p.units + u));
            } else if (u < p.n_units) {
                w.x = __ldg(p.units + u);
                if (u + 1 < p.n_units) w.y = __ldg(p.units + u + 1);
                if (u + 2 < p.n_units) w.z = __ldg(p.units + u + 2);
            }
            pf_w[q] = w;
        }
        {   // look-ahead: the first two units of the next chunk
            const uint64_t u = u0 + 4u * SEGS;
            uint2 w = make_uint2(0u, 0u);
            if (u + 2 <= p.n_units) w = __ldg(reinterpret_cast<const uint2 *>(p.units + u));
            else if (u < p.n_units) w.x = __ldg(p.units + u);
            pf_la = w;
        }
        const uint64_t j1 = g0 + nseg - 1;   // gap element = entry offset of the NEXT chunk
        pf_gw1 = __ldg(p.gap_words + (j1 >> 3));
        uint32_t sh0 = 0;
        if (g0 == 0) {
            pf_gw0 = p.entry_word ? (__ldg(p.entry_word) >> 28) : p.first_bit;
        } else {
            pf_gw0 = __ldg(p.gap_words + ((g0 - 1) >> 3));
            sh0 = 4u * (uint32_t)((g0 - 1) & 7);
        }
        pf_sh = sh0 | (4u * (uint32_t)(j1 & 7)) << 8;
        pf_bits = nseg * kSegmentBits;
    };
    prefetch(s_tileid[0]);

    for (uint32_t k = 0;; ++k) {
        const uint32_t b = k & 1u;
        const uint64_t tile = s_tileid[k % 3u];
        const bool valid = tile < p.n_tiles;
        uint32_t cnt = 0, incl = 0;
        const bool tr = p.trace && blockIdx.x == kTraceCta && k < kTraceSteps && t == 0;
        if (tr) p.trace[k * 8 + 0] = gtime_ns();
        if (valid) {
            // ---------------- decode this thread's chunk (units were prefetched) ----------------
            if (pf_bits) {
                const uint32_t pf_entry = (pf_gw0 >> (pf_sh & 31u)) & 15u;
                const uint32_t pf_stop = pf_bits + ((pf_gw1 >> (pf_sh >> 8)) & 15u);
#pragma unroll
                for (int q = 0; q < SEGS; ++q) {
                    col[(4 * q + 0) * NT] = pf_w[q].x;
                    col[(4 * q + 1) * NT] = pf_w[q].y;
                    col[(4 * q + 2) * NT] = pf_w[q].z;
                    col[(4 * q + 3) * NT] = pf_w[q].w;
                }
                col[(4 * SEGS + 0) * NT] = pf_la.x;
                col[(4 * SEGS + 1) * NT] = pf_la.y;
                if (tr) p.trace[k * 8 + 7] = gtime_ns();
                cnt = (HD_WHATIF & 4) ? 43u : decode_chunk<NT, LONG>(col, pf_entry, pf_stop, T, stage_abs + b * stage_buf_bytes + t * 4u, nullptr);
            }
            // ---------------- warp-wide inclusive scan of the counts ----------------
            incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += v;
            }
            // this warp's TMA store of two steps ago read the tile buffer that is written again next step
            if (store_pending) { tma_bulk_wait_read0(); store_pending = false; }
            if (lane == 31) {
                s_wtot[b][warp] = incl;
                // the last warp to finish publishes the tile aggregate right away (tile 0: it is its own prefix)
                atomicAdd(&s_sum[b], incl);
                __threadfence_block();
                if (atomicAdd(&s_arrived[b], 1u) == kNW - 1) {
                    const uint32_t total = atomicExch(&s_sum[b], 0u);
                    s_arrived[b] = 0;
                    red_max_gpu_u64(p.tile_status + tile * kStatusStride, (tile == 0 ? kFlagPrefix : kFlagAggregate) | (uint64_t)total);
                }
            }
            __threadfence_block();
            if (tr) p.trace[k * 8 + 1] = gtime_ns();
            named_bar_arrive(kBarFull + b, NT + 32);
        }
        if (prev_valid) {
            if (p.tile_buffers == 1u) {
                // single output tile: every warp's TMA store of the previous step must have finished reading it
                // before any warp compacts into it again
                if (store_pending) { tma_bulk_wait_read0(); store_pending = false; }
                named_bar_sync(kBarDecode, NT);
            }
            // wait for the scan warp's prefix of tile k-1 (computed while we decoded tile k): a flag, not a
            // barrier, so that with two output tiles no decode warp ever waits for another decode warp
            mbar_wait(&s_ready[b ^ 1u], ((k - 1u) >> 1) & 1u);
        }
        if (k == 0) {
            while (ld_acquire_cta_u32(&s_t1_ready) == 0) __nanosleep(100);   // the scan warp's second ticket
        }
        if (valid) prefetch(s_tileid[(k + 1u) % 3u]);   // published with s_t1_ready / s_ready above
        if (prev_valid && !(HD_WHATIF & 1)) {
            // ---------------- finish tile k-1: compact this WARP's symbols and flush them ----------------
            const uint32_t pb = b ^ 1u;
            const uint64_t base = s_base[pb];
            const uint32_t woff = s_wtot[pb][warp];                            // this warp's offset in the tile
            const uint32_t wtot = __shfl_sync(0xffffffffu, prev_incl, 31);     // this warp's symbols
            const uint32_t excl_t = woff + prev_incl - prev_cnt;
            // clamp to n_symbols (pad bits decode to garbage: reference clamps with
            // out_pos < next_out_pos, cuhd_gpu_decoder.cu:101-105)
            const uint64_t my_off = base + excl_t;
            uint32_t c = prev_cnt;
            if (my_off >= n_symbols) c = 0;
            else if (my_off + c > n_symbols) c = (uint32_t)(n_symbols - my_off);
            const uint32_t A = (uint32_t)((reinterpret_cast<uintptr_t>(out) + base) & 15u);
            uint8_t *tile_buf = s_tile + (p.tile_buffers == 2u ? pb * p.tile_bytes : 0u);
            if (c && !(HD_WHATIF & 16)) compact_to_tile(reinterpret_cast<const uint32_t *>(s_stage + pb * stage_buf_bytes) + t, NT, c, tile_buf, A + excl_t);
            if (!(HD_WHATIF & 128)) fence_proxy_async_smem();
            __syncwarp();
            // flush: the warp's span is contiguous in the tile and in the output, and both have the same 16-byte
            // phase -> one TMA bulk store for the aligned body, lanes 0-15 / 16-31 for head / tail bytes
            const uint64_t wbase = base + woff;
            uint64_t lim = wbase >= n_symbols ? 0 : n_symbols - wbase;
            const uint32_t total = wtot < lim ? wtot : (uint32_t)lim;
            const uint8_t *sp = tile_buf + A + woff;
            uint8_t *gp = out + wbase;
            uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(gp) & 15u)) & 15u;
            if (head > total) head = total;
            const uint32_t body = (total - head) & ~15u;
            const uint32_t tail = total - head - body;
            if (lane == 0 && body && !(HD_WHATIF & 32)) {
                tma_bulk_store(gp + head, sp + head, body);
                tma_bulk_commit();
                store_pending = true;
            }
            if (HD_WHATIF & 64) {
            } else if (lane < 16) {
                if (lane < head) gp[lane] = sp[lane];
            } else {
                const uint32_t i = lane - 16;
                if (i < tail) gp[head + body + i] = sp[head + body + i];
            }
        }
        if (tr) p.trace[k * 8 + 3] = gtime_ns();
        if (!valid) break;
        prev_valid = true;
        prev_cnt = cnt;
        prev_incl = incl;
    }
    if (store_pending) tma_bulk_wait0();
}

// =========================================================================
// K1: synchronisation = building the gap array of a stream whose codeword
// boundaries are unknown (replaces phases 1+2 of CUHD, cuhd_gpu_decoder.cu:
// 145-327, and its host loop :461-495).
//
// K1a  one thread per CHUNK of 16 segments.  The thread starts parsing 4
//      segments (512 bits) BEFORE its chunk from bit 0 of that halo: by the time
//      it reaches its own chunk the parse has, with high probability, merged
//      with the true one (mean self-synchronisation distance of generate.cpp
//      data at redundancy 0.5: 324 bits), so every landing it records inside its
//      chunk is final.  One full-speed pass with all lanes busy; no block-level
//      fixed-point iteration.  It also records the offset at which its parse
//      entered the chunk (cur_entry).
// K1b  one thread per group of 32 chunks walks the group's seams: if the
//      predecessor's landing into a chunk differs from the entry the chunk's
//      landings were parsed from, re-parse the chunk head from the exact entry
//      until the parses merge (sync_repair_chunk).  Sparse and short.  Rounds
//      repeat until no group exit changes: a code that never self-synchronises
//      degrades to a sequential chain instead of producing garbage.
//
// K1a is persistent: one CTA per SM, every WARP pulls tiles of 32 chunks from an
// atomic counter and stages its tile's units with 16-byte cp.async into a
// private, padded (one pad unit per 16: conflict-free LDS.128 per quarter warp)
// shared-memory region -- coalesced global reads, no block barrier at all.
// =========================================================================
// loads the four units of segment s (zeros past the end of the stream)
struct GlobalSegLoader {
    const uint32_t *units;
    uint64_t n_units;
    __device__ __forceinline__ Seg4 operator()(uint64_t s) const {
        Seg4 r;
        if (s < (n_units >> 2)) {
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(units) + s);
            r.x = w.x; r.y = w.y; r.z = w.z; r.w = w.w;
            return r;
        }
        r.x = r.y = r.z = r.w = 0;
        const uint64_t u = s * 4;
        if (u < n_units) r.x = __ldg(units + u);
        if (u + 1 < n_units) r.y = __ldg(units + u + 1);
        if (u + 2 < n_units) r.z = __ldg(units + u + 2);
        return r;
    }
};

__device__ __forceinline__ void sync_stage_tables(const TableDev &tb, uint8_t *smem, SyncLut &L) {
    const uint32_t lm = tb.lut_bits;
    const uint32_t syn_bytes = 2u << lm;
    uint16_t *s_syn = reinterpret_cast<uint16_t *>(smem);
    uint8_t *s_order = smem + ((syn_bytes + 15u) & ~15u);
    LongTables *s_long = reinterpret_cast<LongTables *>(s_order + 256);
    if (syn_bytes >= 16) {
        const uint4 *src = reinterpret_cast<const uint4 *>(tb.blob + tb.off_sync);
        uint4 *dst = reinterpret_cast<uint4 *>(s_syn);
        for (uint32_t i = threadIdx.x; i < syn_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
        for (uint32_t i = threadIdx.x; i < syn_bytes / 2; i += blockDim.x)
            s_syn[i] = __ldg(reinterpret_cast<const uint16_t *>(tb.blob + tb.off_sync) + i);
    }
    if (tb.has_long) {
        for (uint32_t i = threadIdx.x; i < 64; i += blockDim.x)
            reinterpret_cast<uint32_t *>(s_order)[i] = __ldg(reinterpret_cast<const uint32_t *>(tb.blob + tb.off_order) + i);
        for (uint32_t i = threadIdx.x; i < sizeof(LongTables) / 4; i += blockDim.x)
            reinterpret_cast<uint32_t *>(s_long)[i] = __ldg(reinterpret_cast<const uint32_t *>(tb.blob + tb.off_long) + i);
    }
    L.lut = s_syn;
    L.idx_shift = 32 - lm;
    L.lm = lm;
    L.lt = s_long;
    L.order = s_order;
}
__device__ __forceinline__ uint32_t sync_tables_bytes(uint32_t lut_bits) {
    return (((2u << lut_bits) + 15u) & ~15u) + 256u + (((uint32_t)sizeof(LongTables) + 15u) & ~15u);
}

// 16-byte async copy global -> shared (LDGSTS), zero-filling the bytes past src_bytes
__device__ __forceinline__ void cp_async16_zfill(void *sdst, const void *gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// loads the four units of (global) segment s from a warp's staged, padded tile that starts at segment q0
struct TileSegLoader {
    const uint4 *tile;
    int64_t q0;
    __device__ __forceinline__ Seg4 operator()(uint64_t s) const {
        const uint32_t i = (uint32_t)((int64_t)s - q0);
        const uint4 w = tile[i + (i >> 4)];
        Seg4 r;
        r.x = w.x; r.y = w.y; r.z = w.z; r.w = w.w;
        return r;
    }
};

template <bool EXACT, bool LONG>
__global__ void __launch_bounds__(kSyncWarps * 32, 1) hd_sync_parse_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    SyncLut L;
    sync_stage_tables(p.table, smem, L);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *tile = reinterpret_cast<uint4 *>(smem + sync_tables_bytes(p.table.lut_bits) + (size_t)warp * kSyncTileBytes);
    const uint64_t n_units16 = (p.n_units + 3) >> 2;   // 16-byte units that hold stream data (last one may be partial)

    // ---- the seam between two warp tiles.  Lane 1 of tile t entered its chunk with the exit lane 0 computed for the
    // chunk before it, which is owned (and possibly repaired) by tile t-1.  Every tile PUBLISHES the exit of its last
    // chunk; the warp compares the published exit of tile t-1 with what tile t assumed ONE TILE LATER, when it has long
    // been there (tiles t-1 and t are parsed at the same time by two warps: checking right away would mean waiting).
    // The rare mismatch (the chunk before the tile had not synchronised after 20 segments) is repaired from global
    // memory by one lane, walking forward until the parses merge; a walk that runs through the whole tile changes an
    // exit that is already published: `changed` is counted and the host runs K1b rounds.
    const uint32_t kW = kSyncChunkSegs / 8;
    uint32_t pend_tile = 0, pend_assumed = 0;   // tile whose first seam is still to be checked (0: none)
    auto check_seam = [&]() {
        if (pend_tile == 0) return;
        uint32_t st = 0;
        if (lane == 0) {
            for (uint32_t polls = 0; polls < kExitWaitPolls; ++polls) {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(st) : "l"(p.tile_exit + pend_tile - 1) : "memory");
                if (st & kExitPublished) break;
                __nanosleep(200);
            }
            if (!(st & kExitPublished)) {
                atomicAdd(p.changed, 1u);   // predecessor not seen in time (never expected): the K1b rounds settle it
            } else if ((st & 15u) != pend_assumed) {
                const GlobalSegLoader gld{p.units, p.n_units};
                const uint64_t c0 = (uint64_t)pend_tile * kSyncTileChunks + 1;
                uint64_t c1 = c0 + kSyncTileChunks;
                if (c1 > p.n_chunks) c1 = p.n_chunks;
                uint32_t xin = st & 15u;
                bool exit_changed = false;
                for (uint64_t c = c0; c < c1; ++c) {
                    uint32_t words[2] = {p.gap_words[c * kW], p.gap_words[c * kW + 1]};
                    const bool merged = sync_repair_chunk<EXACT, LONG>(gld, c, xin, L, words);
                    p.gap_words[c * kW] = words[0];
                    p.gap_words[c * kW + 1] = words[1];
                    p.cur_entry[c] = (uint8_t)xin;
                    exit_changed = !merged;
                    if (merged) break;
                    xin = words[1] >> 28;
                }
                if (exit_changed && c1 - c0 == kSyncTileChunks) atomicAdd(p.changed, 1u);
            }
        }
        __syncwarp();
        pend_tile = 0;
    };

    for (;;) {
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(p.tile_counter, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= p.n_tiles) break;
        // ---- stage [halo | 512 segments | 1 look-ahead segment] of this warp tile ----
        const uint64_t chunk0 = (uint64_t)ticket * kSyncTileChunks;   // lane 0's chunk (owned by the previous tile)
        const int64_t q0 = (int64_t)(chunk0 * kSyncChunkSegs) - kSyncHaloSegs;
        __syncwarp();   // every lane is done reading the previous tile
        for (uint32_t i = lane; i < kSyncTileUnits16; i += 32) {
            const int64_t q = q0 + i;
            uint32_t bytes = 0;
            if (q >= 0 && (uint64_t)q < n_units16) {
                const uint64_t rem = p.n_units - (uint64_t)q * 4;
                bytes = rem >= 4 ? 16u : (uint32_t)rem * 4u;
            }
            // (a source address is needed even when nothing is copied)
            cp_async16_zfill(tile + i + (i >> 4), p.units + (bytes ? (uint64_t)q * 4 : 0), bytes);
        }
        cp_async_commit_wait_all();
        __syncwarp();

        // ---- parse: 4 halo segments from bit 0, then the 16 segments of the chunk ----
        const uint64_t chunk = chunk0 + lane;
        const bool active = chunk < p.n_chunks;
        const bool owner = active && (lane > 0 || ticket == 0);
        uint32_t words[2] = {0u, 0u};
        uint32_t entry = 0;
        if (active && !(HD_WHATIF & 256)) {
            uint32_t i = kSyncChunkSegs * lane;   // unit index (within the staged tile) of the halo start
            uint32_t pos = 0;
            if (chunk == 0) {
                pos = p.first_bit;                // the stream's first chunk starts at a known offset: no halo
                i += kSyncHaloSegs;
            } else {
#pragma unroll 1
                for (uint32_t h = 0; h < kSyncHaloSegs; ++h, ++i) {
                    const uint4 w = tile[i + (i >> 4)];
                    const uint32_t n = tile[i + 1 + ((i + 1) >> 4)].x;
                    pos = sync_segment<EXACT, LONG>(w.x, w.y, w.z, w.w, n, pos, L);
                }
            }
            entry = pos;
#pragma unroll 1
            for (uint32_t g = 0; g < kSyncChunkSegs / 8; ++g) {
                uint32_t word = 0;
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k, ++i) {
                    const uint4 w = tile[i + (i >> 4)];
                    const uint32_t n = tile[i + 1 + ((i + 1) >> 4)].x;
                    pos = sync_segment<EXACT, LONG>(w.x, w.y, w.z, w.w, n, pos, L);
                    word |= pos << (4u * k);
                }
                words[g] = word;
            }
        }
        // ---- seams inside the warp tile: a lane whose halo parse had not merged yet (~20 % at the reference's
        // redundancy 0.5) re-parses the head of its chunk from its left neighbour's landing, from the staged
        // tile; repeated until no lane changes (a parse that fails to merge inside a chunk moves on one lane per
        // round).  After this the tile is consistent given the entry of its FIRST owned chunk; check_seam() looks at that one.
        const TileSegLoader tld{tile, q0};
        for (;;) {
            const uint32_t left = __shfl_up_sync(0xffffffffu, words[1] >> 28, 1);
            const bool redo = active && lane > 0 && left != entry && !(HD_WHATIF & 512);
            if (!__any_sync(0xffffffffu, redo)) break;
            if (redo) {
                sync_repair_chunk<EXACT, LONG>(tld, chunk, left, L, words);
                entry = left;
            }
        }
        if (owner) {
            p.cur_entry[chunk] = (uint8_t)entry;
            *reinterpret_cast<uint2 *>(p.gap_words + chunk * (kSyncChunkSegs / 8)) = make_uint2(words[0], words[1]);
        }
        if (lane == 31 && active) {   // (a tile without a 32nd chunk is the last one: nobody reads its exit)
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.tile_exit + ticket), "r"(kExitPublished | (words[1] >> 28)) : "memory");
        }
        const uint32_t assumed = __shfl_sync(0xffffffffu, words[1] >> 28, 0);   // lane 0's exit = what lane 1 was entered with
        check_seam();   // of the tile this warp parsed BEFORE this one
        if (ticket > 0) { pend_tile = ticket; pend_assumed = assumed; }
        __syncwarp();   // the repair above wrote this warp's earlier results; keep the warp together before restaging
    }
    check_seam();
}

// loads segment s from a thread-private shared-memory copy of one chunk that starts at segment seg0
struct SlotSegLoader {
    const Seg4 *slot;
    uint64_t seg0;
    __device__ __forceinline__ Seg4 operator()(uint64_t s) const { return slot[s - seg0]; }
};

// K1b: seam check + repair.  One thread owns a GROUP of kSyncWalkChunks (= one
// K1a warp tile) consecutive chunks and walks its seams in order: where the
// landing of the previous chunk differs from the entry the chunk's landings
// were parsed from, it re-parses the chunk head from the exact entry until the
// parses merge (sync_repair_chunk), carrying the exact exit into the next chunk
// (Gauss-Seidel inside the group; no two threads ever write the same chunk).
// `changed` counts the groups whose EXIT changed: the host launches rounds until
// it stays zero.  One round suffices unless a parse fails to synchronise within a
// whole group (64 Kbit); a code that never synchronises (all codewords the same
// length not dividing 128) advances one group per round -- the true parse of
// such a code is inherently sequential; this is the graceful worst case, the
// reference emits garbage here.
constexpr uint32_t kSyncWalkChunks = kSyncTileChunks;   // group g = the chunks owned by K1a tile g: [31g+1, 31g+32)

template <bool EXACT, bool LONG>
__global__ void __launch_bounds__(128) hd_sync_walk_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    SyncLut L;
    sync_stage_tables(p.table, smem, L);
    __syncthreads();
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t c0 = g * kSyncWalkChunks + 1;   // chunk 0 starts at a known offset and is never repaired
    if (c0 >= p.n_chunks) return;
    uint64_t c1 = c0 + kSyncWalkChunks;
    if (c1 > p.n_chunks) c1 = p.n_chunks;
    const uint32_t kW = kSyncChunkSegs / 8;
    const GlobalSegLoader ld{p.units, p.n_units};
    // a repair first pulls the whole chunk (+2 look-ahead segments) into this thread's shared-memory slot with
    // back-to-back 128-bit loads, so that the parse pays ONE global-memory latency instead of one per segment
    Seg4 *slot = reinterpret_cast<Seg4 *>(smem + sync_tables_bytes(p.table.lut_bits)) + (size_t)threadIdx.x * (kSyncChunkSegs + 2);
    uint32_t xin = p.gap_words[c0 * kW - 1] >> 28;
    bool exit_changed = false;
    for (uint64_t c = c0; c < c1; ++c) {
        uint32_t words[2] = {p.gap_words[c * kW], p.gap_words[c * kW + 1]};
        exit_changed = false;
        if (xin != p.cur_entry[c]) {
            const uint64_t seg0 = c * kSyncChunkSegs;
#pragma unroll
            for (uint32_t k = 0; k < kSyncChunkSegs + 2; ++k) slot[k] = ld(seg0 + k);
            const SlotSegLoader sld{slot, seg0};
            const bool merged = sync_repair_chunk<EXACT, LONG>(sld, c, xin, L, words);
            p.gap_words[c * kW] = words[0];
            p.gap_words[c * kW + 1] = words[1];
            p.cur_entry[c] = (uint8_t)xin;
            exit_changed = !merged;
        }
        // K1a left every group consistent given the entry of its first chunk, and this walk keeps it so: once a
        // chunk's exit is unchanged, nothing after it can change
        if (!exit_changed) break;
        xin = words[1] >> 28;
    }
    if (exit_changed) atomicAdd(p.changed, 1u);
}

}  // namespace hd
#endif  // HD_KERNELS_CUH_
