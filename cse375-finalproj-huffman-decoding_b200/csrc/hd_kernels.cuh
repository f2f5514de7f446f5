// hd_kernels.cuh -- device code of the B200 (sm_100a) Huffman decoder.
//
//   K4  hd_decode_tiles_kernel : see hd_decode.cuh (this file holds the PTX helpers and the look-back it uses)
//   K1a hd_sync_parse_kernel   : speculative subsequence synchronisation: every
//        thread parses a halo + its chunk, seams are repaired inside the warp
//        (replaces phase 1, cuhd_gpu_decoder.cu:145-233)
//   K1b hd_sync_walk_kernel    : seams between warp tiles, walked to the merge
//        (replaces phase 2 + host loop, cuhd_gpu_decoder.cu:235-327, 461-495)
//   K1c hd_sync_map_kernel + hd_sync_chase_kernel : per-tile transfer functions and the chase through them, for codes
//        that never synchronise (no counterpart in the reference: it emits garbage there)
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
    // second-level fallback (codes that do not synchronise within whole tiles: K1c + chase, then K1a once more):
    unsigned long long *tile_map;   // per tile: 16 nibbles, nibble x = exit offset of the tile when it is entered at offset x
    uint8_t *tile_entry;            // per tile: the exact offset at which the true parse enters its first owned chunk
    uint32_t use_tile_entry;        // K1a: take lane 1's entry from tile_entry instead of speculating (no seam checks)
    uint32_t k_one;                 // 1 (opaque to ptxas: keeps the window slide of sync_run on the FMA pipe)
};
constexpr uint32_t kExitPublished = 0x100u;
constexpr uint32_t kExitWaitPolls = 1u << 16;   // bounded wait for a predecessor tile (then: flag `changed`, K1b rounds decide)

// ---- helpers shared with K4 (hd_decode.cuh) ------------------------------------------------------------------
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
// The wait is bounded (kLookbackMaxPolls rounds of ~1 us): a predecessor that never publishes -- a CTA of the
// persistent grid that is not running, which the ticket scheme rules out, or a corrupted workspace -- raises *stuck
// instead of hanging the GPU; the decode then finishes with wrong offsets and the entry points that synchronise
// report HD_ERR_UNRESOLVED.
constexpr uint32_t kLookbackMaxPolls = 1u << 22;
__device__ __forceinline__ uint64_t lookback_wide(const uint64_t *status, uint64_t tile, uint32_t lane, uint32_t *stuck) {
    uint64_t excl = 0;
    uint32_t n_polls = 0;
    int64_t hi = (int64_t)tile - 1;
    for (;;) {
        const int64_t first = hi - 8 * (int64_t)lane;
        if (++n_polls > kLookbackMaxPolls) {
            if (lane == 0) atomicExch(stuck, 1u);
            break;
        }
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
    return excl;
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
// K1a, EXACT = false: [padding < 2^lm][u8 jump[2^lm], aligned to its size][order][LongTables] -- sized for the worst padding
__host__ __device__ __forceinline__ uint32_t sync_jump8_tables_bytes(uint32_t lut_bits) {
    return (1u << lut_bits) + (((1u << lut_bits) + 15u) & ~15u) + 256u + (((uint32_t)sizeof(LongTables) + 15u) & ~15u);
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

// loads segment s of the chunk that starts at segment seg0 from the owning lane's region of the staged tile
struct RegionSegLoader {
    const uint4 *region;
    uint64_t seg0;
    __device__ __forceinline__ Seg4 operator()(uint64_t s) const {
        const uint4 w = region[(uint32_t)(s - seg0)];
        Seg4 r;
        r.x = w.x; r.y = w.y; r.z = w.z; r.w = w.w;
        return r;
    }
};

template <bool EXACT, bool LONG>
__global__ void __launch_bounds__(kSyncWarps * 32, 1) hd_sync_parse_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    SyncLut L;
    uint32_t tiles_off;
    if constexpr (!EXACT && !(HD_WHATIF & 2048)) {
        // The decode path's landing rule (EXACT = false) needs only the jump of an entry: a BYTE table, placed in the
        // dynamic window at an address that is a multiple of its size, so that the funnel shift that extracts the
        // window's index bits also ORs in the table's address (the base goes in as the shift's upper word): the load
        // needs no address arithmetic at all, and four entries share a bank word instead of two.
        const uint32_t lm = p.table.lut_bits, tsz = 1u << lm;
        const uint32_t abs0 = (uint32_t)__cvta_generic_to_shared(smem);
        const uint32_t pad = (0u - abs0) & (tsz - 1u);   // (< tsz; the host sized the window for tsz)
        uint8_t *s_jump8 = smem + pad;
        uint8_t *s_order = s_jump8 + ((tsz + 15u) & ~15u);
        LongTables *s_long = reinterpret_cast<LongTables *>(s_order + 256);
        const uint16_t *src = reinterpret_cast<const uint16_t *>(p.table.blob + p.table.off_sync);
        for (uint32_t i = threadIdx.x; i < tsz; i += blockDim.x) s_jump8[i] = (uint8_t)__ldg(src + i);
        if (p.table.has_long) {
            for (uint32_t i = threadIdx.x; i < 64; i += blockDim.x)
                reinterpret_cast<uint32_t *>(s_order)[i] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_order) + i);
            for (uint32_t i = threadIdx.x; i < sizeof(LongTables) / 4; i += blockDim.x)
                reinterpret_cast<uint32_t *>(s_long)[i] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_long) + i);
        }
        L.lut = nullptr;
        L.idx_shift = 32 - lm;
        L.lm = lm;
        L.lt = s_long;
        L.order = s_order;
        L.use8 = true;
        L.jump8_hi = (abs0 + pad) >> lm;
        tiles_off = pad + sync_jump8_tables_bytes(lm) - tsz;
    } else {
        sync_stage_tables(p.table, smem, L);
        tiles_off = sync_tables_bytes(p.table.lut_bits);
    }
    L.one = p.k_one;
    __syncthreads();
    // the decode kernel may be launched as a programmatic dependent of this one: let its CTAs be placed as soon as SMs
    // free up (they wait for this grid to complete before they read the gap words)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *tile = reinterpret_cast<uint4 *>(smem + tiles_off + (size_t)warp * kSyncTileBytes);
    const uint32_t land_abs = (uint32_t)__cvta_generic_to_shared(tile + kSyncTilePhysUnits);   // landing bytes [slot][lane]
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
        const bool interior = q0 >= 0 && (uint64_t)q0 + kSyncTileSrcUnits <= (p.n_units >> 2);   // every unit is whole
        if (interior && !(HD_WHATIF & 4096)) {
            // Source unit i = lane + 32 k sits at physical unit sync_phys_unit(lane) + 34 k (32 units = two regions of
            // 17), and it is the first unit of a chunk exactly for lanes 4 and 20: the whole staging is 17 copies at
            // constant offsets from two per-lane addresses, fully unrolled (the generic loop below spends ~28
            // instructions per copy on the layout arithmetic).
            static_assert(kSyncHaloSegs == 4 && kSyncChunkSegs == 16 && kSyncRegionUnits == 17, "staging offsets");
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + sync_phys_unit(lane));
            const uint4 *src = reinterpret_cast<const uint4 *>(p.units) + q0 + lane;
            const bool first = lane == 4 || lane == 20;
#pragma unroll
            for (uint32_t k = 0; k < kSyncTileSrcUnits / 32; ++k) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 34u * 16u * k), "l"(src + 32u * k) : "memory");
                if (first) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 34u * 16u * k - 16u), "l"(src + 32u * k) : "memory");
            }
            if (lane < kSyncTileSrcUnits % 32) {   // units 512..516: the look-ahead segment (+ the 4 units a repair may read)
                const uint32_t k = kSyncTileSrcUnits / 32;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 34u * 16u * k), "l"(src + 32u * k) : "memory");
                if (first) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 34u * 16u * k - 16u), "l"(src + 32u * k) : "memory");
            }
        } else {
#pragma unroll 1
        for (uint32_t i = lane; i < kSyncTileSrcUnits; i += 32) {
            const int64_t q = q0 + i;
            uint32_t bytes = 16;
            if (!interior) {
                bytes = 0;
                if (q >= 0 && (uint64_t)q < n_units16) {
                    const uint64_t rem = p.n_units - (uint64_t)q * 4;
                    bytes = rem >= 4 ? 16u : (uint32_t)rem * 4u;
                }
            }
            // (a source address is needed even when nothing is copied)
            const uint32_t *src = p.units + (bytes ? (uint64_t)q * 4 : 0);
            uint4 *dst = tile + sync_phys_unit(i);
            cp_async16_zfill(dst, src, bytes);
            // the first unit of a chunk also goes into the slot behind the previous region (or the tile halo)
            if (i >= kSyncHaloSegs && ((i - kSyncHaloSegs) & (kSyncChunkSegs - 1)) == 0) cp_async16_zfill(dst - 1, src, bytes);
        }
        }
        cp_async_commit_wait_all();
        __syncwarp();

        // ---- parse: 4 halo segments from bit 0, then the 16 segments of the chunk: two flattened runs ----
        const uint64_t chunk = chunk0 + lane;
        const bool active = chunk < p.n_chunks;
        const bool owner = active && (lane > 0 || ticket == 0);
        uint32_t words[2] = {0u, 0u};
        uint32_t entry = 0;
        if (active && !(HD_WHATIF & 256)) {
            const uint32_t *region = reinterpret_cast<const uint32_t *>(tile + kSyncHaloSegs + 1 + kSyncRegionUnits * lane);
            uint32_t pos = p.first_bit;   // the stream's first chunk starts at a known offset: no halo
            if (chunk != 0) {
                // the halo = the 4 units before the region (lane 0: the tile halo), followed by a copy of the chunk's
                // first unit
                pos = sync_run<EXACT, LONG, false>(region - 4 * (kSyncHaloSegs + 1), 0u, kSyncHaloSegs * kSegmentBits, L, 0u, nullptr) -
                      kSyncHaloSegs * kSegmentBits;
            }
            entry = pos;
            sync_run<EXACT, LONG, true>(region, entry, kSyncChunkSegs * kSegmentBits, L, land_abs + 32u + lane, nullptr);
            // landings into segments 1..16 of the chunk (slot k = landing into segment k)
#pragma unroll
            for (uint32_t k = 0; k < kSyncChunkSegs; ++k) {
                uint32_t b;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(land_abs + 32u * (k + 1u) + lane));
                words[k >> 3] |= (b & 15u) << (4u * (k & 7u));
            }
        }
        // ---- seams inside the warp tile: a lane whose halo parse had not merged yet (~20 % at the reference's
        // redundancy 0.5) re-parses the head of its chunk from its left neighbour's landing, from the staged
        // tile; repeated until no lane changes (a parse that fails to merge inside a chunk moves on one lane per
        // round).  After this the tile is consistent given the entry of its FIRST owned chunk; check_seam() looks at that one.
        // (a lane repairs its own chunk: segment seg0 + k is unit k of its region; k = 17 -- read ahead by
        // sync_repair_chunk, never used -- falls on the next region's first unit)
        const RegionSegLoader tld{tile + kSyncHaloSegs + 1 + kSyncRegionUnits * lane, chunk * kSyncChunkSegs};
        for (;;) {
            uint32_t left = __shfl_up_sync(0xffffffffu, words[1] >> 28, 1);
            if (p.use_tile_entry && lane == 1 && ticket > 0) left = p.tile_entry[ticket];   // exact (chased through the tile maps)
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
        if (ticket > 0 && !p.use_tile_entry) { pend_tile = ticket; pend_assumed = assumed; }
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

// =========================================================================
// Second-level fallback for codes that do not synchronise within whole tiles (e.g. all codewords 3 bits: the parse of
// such a code is a sequential chain, and K1b moves it one tile per round).  The chain is turned into a PREFIX problem
// over TRANSFER FUNCTIONS: the offset at which the parse leaves a tile depends only on the offset (0..15) at which it
// enters it, so every tile has a map {0..15} -> {0..15}; the maps of all tiles are computed in parallel (K1c: 16 lanes
// per tile, lane x parses the tile from entry offset x), one thread chases the true entry through them (16 nibbles per
// tile, staged through shared memory), and K1a runs once more with the exact tile entries.  Any prefix code, no rounds.
// =========================================================================
template <bool EXACT, bool LONG>
__global__ void __launch_bounds__(256) hd_sync_map_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    SyncLut L;
    sync_stage_tables(p.table, smem, L);
    __syncthreads();
    const uint64_t half = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // one half warp per tile
    const uint32_t x = threadIdx.x & 15u;
    if (half >= p.n_tiles) return;
    const uint64_t c0 = half * kSyncTileChunks + 1;
    uint64_t c1 = c0 + kSyncTileChunks;
    if (c1 > p.n_chunks) c1 = p.n_chunks;
    uint32_t pos = x;
    if (c0 < c1) {
        const GlobalSegLoader ld{p.units, p.n_units};
        uint64_t s0 = c0 * kSyncChunkSegs, s1 = c1 * kSyncChunkSegs;
        if (s1 > p.n_segments) s1 = p.n_segments;
        Seg4 cur = ld(s0);
        for (uint64_t sg = s0; sg < s1; ++sg) {
            const Seg4 nxt = ld(sg + 1);
            pos = sync_segment<EXACT, LONG>(cur.x, cur.y, cur.z, cur.w, nxt.x, pos, L);
            cur = nxt;
        }
    }
    atomicOr(p.tile_map + half, (unsigned long long)(pos & 15u) << (4u * x));   // (zeroed before launch)
}

// one CTA: thread 0 parses chunk 0 from the stream's known start, then chases the entry through the tile maps
template <bool EXACT, bool LONG>
__global__ void __launch_bounds__(256) hd_sync_chase_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    SyncLut L;
    sync_stage_tables(p.table, smem, L);
    __shared__ unsigned long long s_map[2048];
    __shared__ uint8_t s_entry[2048];
    __shared__ uint32_t s_x;
    __syncthreads();
    if (threadIdx.x == 0) {
        const GlobalSegLoader ld{p.units, p.n_units};
        uint32_t pos = p.first_bit;
        uint64_t s1 = kSyncChunkSegs < p.n_segments ? kSyncChunkSegs : p.n_segments;
        Seg4 cur = ld(0);
        for (uint64_t sg = 0; sg < s1; ++sg) {
            const Seg4 nxt = ld(sg + 1);
            pos = sync_segment<EXACT, LONG>(cur.x, cur.y, cur.z, cur.w, nxt.x, pos, L);
            cur = nxt;
        }
        s_x = pos & 15u;
    }
    for (uint64_t base = 0; base < p.n_tiles; base += 2048) {
        const uint64_t n = p.n_tiles - base < 2048 ? p.n_tiles - base : 2048;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_map[i] = p.tile_map[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t xx = s_x;
            for (uint32_t i = 0; i < n; ++i) {
                s_entry[i] = (uint8_t)xx;
                xx = (uint32_t)(s_map[i] >> (4u * xx)) & 15u;
            }
            s_x = xx;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) p.tile_entry[base + i] = s_entry[i];
    }
}

}  // namespace hd
#endif  // HD_KERNELS_CUH_
