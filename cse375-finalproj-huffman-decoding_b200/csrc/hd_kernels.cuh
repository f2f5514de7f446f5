// hd_kernels.cuh -- device code of the B200 (sm_100a) Huffman decoder.
//
//   K4  hd_decode_write_kernel : gap-array segment decode + single-pass
//        decoupled look-back prescan + decode-and-write, one persistent launch.
//        Replaces gpu_dec_l1_l2 (Huffman_coding_Gap_arrays/decoder/src/
//        decoder.cu:454-730) and phases 3+4 of CUHD (gpuhd/src/
//        cuhd_gpu_decoder.cu:329-420 + thrust scan :498-505).
//   K1a hd_sync_chunks_kernel  : speculative subsequence synchronisation inside
//        a block to a fixed point (replaces phase 1, cuhd_gpu_decoder.cu:145-233)
//   K1b hd_sync_seams_kernel   : block-seam synchronisation to a fixed point
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

// ------------------------------------------------------------ kernel params
constexpr uint64_t kFlagAggregate = 1ull << 62;   // tile aggregate published
constexpr uint64_t kFlagPrefix = 2ull << 62;      // inclusive prefix published
constexpr uint64_t kFlagMask = 3ull << 62;

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
    uint64_t *tile_status;       // n_tiles words, zeroed before launch
    uint32_t *tile_counter;      // zeroed before launch
    uint64_t n_tiles;
    uint64_t *count_out;         // optional: total symbols written
    uint32_t stage_words;        // staging rows (one 32-bit word per thread each)
    uint32_t tile_bytes;         // output tile bytes (multiple of 16)
};

struct SyncParams {
    const uint32_t *units;
    uint64_t n_units;            // units that may be READ
    uint64_t n_segments;
    uint64_t n_chunks;           // ceil(n_segments / kChunkSegs)
    uint64_t n_blocks;           // ceil(n_chunks / kSyncThreads)
    uint32_t first_bit;
    TableDev table;
    uint32_t *gap_words;         // out: one word per chunk
    uint8_t *blk_entry;          // per sync block: entry offset its chunk 0 assumed
    uint32_t *changed;           // per-round "some block exit changed" counter
};

// =========================================================================
// K4: decode-and-write.  One thread per CHUNK of SEGS consecutive 128-bit
// segments, kTileThreads chunks per tile, persistent CTAs pulling tile ids from
// an atomic counter so that the look-back never waits on a tile that is not
// running.  Per tile:
//   load    each thread copies its chunk's units (128-bit loads) into its column
//           of the transposed input tile (word w of thread t at [w][t])
//   decode  decode_chunk(): flattened loop over the whole chunk, multi-symbol LUT
//           in smem, symbols byte-stored into the thread's staging column
//   scan    warp shuffles + 8 warp totals; warp 0 runs the decoupled look-back
//   compact staging columns -> contiguous output tile (at the 16-byte phase of
//           the global destination)
//   flush   one TMA bulk store (UBLKCP) for the 16-byte aligned body, scalar
//           stores for head/tail
// =========================================================================
__device__ __forceinline__ uint32_t gap_nibble(const uint32_t *gap_words, uint64_t j) {
    return (__ldg(gap_words + (j >> 3)) >> (4u * (uint32_t)(j & 7))) & 15u;
}

template <int SEGS, bool LONG>
__global__ void __launch_bounds__(kTileThreads) hd_decode_write_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    const uint32_t lm = p.table.lut_bits;
    const uint32_t lut_bytes = 4u << lm;
    constexpr uint32_t kInRows = 4 * SEGS + 3;

    uint32_t *s_lut = reinterpret_cast<uint32_t *>(smem);
    uint8_t *s_order = smem + lut_bytes;
    LongTables *s_long = reinterpret_cast<LongTables *>(s_order + 256);
    uint32_t off = lut_bytes + (LONG ? 256u + ((uint32_t)(sizeof(LongTables) + 15) & ~15u) : 0u);
    uint32_t *s_in = reinterpret_cast<uint32_t *>(smem + off);
    off += kInRows * kTileThreads * 4u;
    // the staging rows must start on a 1 KiB boundary of the shared window (see decode_chunk)
    const uint32_t smem_abs = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t stage_abs = (smem_abs + off + 1023u) & ~1023u;
    uint32_t *s_stage = reinterpret_cast<uint32_t *>(smem + (stage_abs - smem_abs));
    uint8_t *s_tile = reinterpret_cast<uint8_t *>(s_stage + (size_t)p.stage_words * kTileThreads);

    __shared__ uint32_t s_warp_tot[kTileThreads / 32];
    __shared__ uint32_t s_tile_id[2];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_total;

    // ---- stage the tables once per CTA (persistent) ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.table.blob);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t i = t; i < lut_bytes / 16; i += kTileThreads) dst[i] = __ldg(src + i);
        if (LONG) {
            if (t < 64) reinterpret_cast<uint32_t *>(s_order)[t] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_order) + t);
            if (t < sizeof(LongTables) / 4) reinterpret_cast<uint32_t *>(s_long)[t] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_long) + t);
        }
    }
    if (t == 0) s_tile_id[0] = atomicAdd(p.tile_counter, 1u);
    __syncthreads();

    DecodeTables T;
    T.lut = s_lut;
    T.idx_shift = 32 - lm;
    T.lut_bits = lm;
    T.len_tab = nullptr;
    T.lt = s_long;
    T.order = s_order;
    uint32_t par = 0;
    bool store_pending = false;
    uint32_t *const col = s_in + t;

    for (;;) {
        const uint64_t tile = s_tile_id[par];
        if (tile >= p.n_tiles) break;
        if (t == 0) s_tile_id[par ^ 1] = atomicAdd(p.tile_counter, 1u);   // prefetch next id

        // ---------------- load + decode this thread's chunk ----------------
        const uint64_t g0 = (tile * kTileThreads + t) * SEGS;
        uint32_t cnt = 0;
        if (g0 < p.n_segments) {
            const uint64_t left = p.n_segments - g0;
            const uint32_t nseg = left < (uint64_t)SEGS ? (uint32_t)left : (uint32_t)SEGS;
            const uint64_t u0 = g0 * 4;
#pragma unroll
            for (int k = 0; k < SEGS; ++k) {
                const uint64_t u = u0 + 4u * k;
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (u + 4 <= p.n_units) {
                    w = __ldg(reinterpret_cast<const uint4 *>(p.units + u));
                } else if (u < p.n_units) {
                    w.x = __ldg(p.units + u);
                    if (u + 1 < p.n_units) w.y = __ldg(p.units + u + 1);
                    if (u + 2 < p.n_units) w.z = __ldg(p.units + u + 2);
                }
                col[(4 * k + 0) * kTileThreads] = w.x;
                col[(4 * k + 1) * kTileThreads] = w.y;
                col[(4 * k + 2) * kTileThreads] = w.z;
                col[(4 * k + 3) * kTileThreads] = w.w;
            }
            {   // look-ahead: the first two units of the next chunk
                const uint64_t u = u0 + 4u * SEGS;
                uint2 w = make_uint2(0u, 0u);
                if (u + 2 <= p.n_units) w = __ldg(reinterpret_cast<const uint2 *>(p.units + u));
                else if (u < p.n_units) w.x = __ldg(p.units + u);
                col[(4 * SEGS + 0) * kTileThreads] = w.x;
                col[(4 * SEGS + 1) * kTileThreads] = w.y;
            }
            const uint32_t entry = g0 == 0 ? (p.entry_word ? (__ldg(p.entry_word) >> 28) : p.first_bit)
                                           : gap_nibble(p.gap_words, g0 - 1);
            const uint32_t stop = nseg * kSegmentBits + gap_nibble(p.gap_words, g0 + nseg - 1);
            cnt = decode_chunk<LONG>(col, entry, stop, T, stage_abs + t * 4u, nullptr);
        }

        // ---------------- block-wide exclusive scan of counts ----------------
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += v;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();   // (B)

        // ---------------- decoupled look-back (warp 0) ----------------
        if (warp == 0) {
            uint32_t wt = lane < kTileThreads / 32 ? s_warp_tot[lane] : 0;
            uint32_t wi = wt;
#pragma unroll
            for (int d = 1; d < kTileThreads / 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += v;
            }
            if (lane < kTileThreads / 32) s_warp_tot[lane] = wi - wt;   // exclusive warp offsets
            const uint64_t agg = __shfl_sync(0xffffffffu, wi, kTileThreads / 32 - 1);
            if (lane == 0) st_relaxed_gpu_u64(p.tile_status + tile, (tile == 0 ? kFlagPrefix : kFlagAggregate) | agg);
            uint64_t excl = 0;
            if (tile > 0) {
                int64_t look = (int64_t)tile - 1;
                for (;;) {
                    const int64_t idx = look - lane;
                    uint64_t v = kFlagPrefix;   // virtual tile -1: prefix 0
                    if (idx >= 0) {
                        do { v = ld_relaxed_gpu_u64(p.tile_status + idx); } while ((v & kFlagMask) == 0);
                    }
                    const uint32_t pm = __ballot_sync(0xffffffffu, (v & kFlagMask) == kFlagPrefix);
                    // lanes at or before the nearest prefix contribute
                    const uint32_t first = pm ? (uint32_t)__ffs(pm) - 1 : 32u;
                    uint64_t c = (lane <= first) ? (v & ~kFlagMask) : 0ull;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                    excl += c;
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) st_relaxed_gpu_u64(p.tile_status + tile, kFlagPrefix | (excl + agg));
            }
            if (lane == 0) {
                s_base = excl;
                s_total = (uint32_t)agg;
                // the previous tile's TMA store must have finished READING s_tile
                if (store_pending) tma_bulk_wait_read0();
                if (tile + 1 == p.n_tiles) {
                    uint64_t tot = excl + agg;
                    tot = tot < p.n_symbols ? tot : p.n_symbols;
                    if (p.count_out) *p.count_out = tot;
                    if (p.shard_info) {
                        p.shard_info[0] = tot;
                        p.shard_info[1] = p.entry_word ? (__ldg(p.entry_word) >> 28) : p.first_bit;
                        p.shard_info[2] = gap_nibble(p.gap_words, p.n_segments - 1);   // where the parse leaves the shard
                    }
                }
            }
        }
        __syncthreads();   // (C)

        // ---------------- compact staging -> output tile ----------------
        const uint64_t base = s_base;
        const uint32_t excl_t = s_warp_tot[warp] + incl - cnt;
        // clamp to n_symbols (pad bits decode to garbage: reference clamps with
        // out_pos < next_out_pos, cuhd_gpu_decoder.cu:101-105)
        uint64_t my_off = base + excl_t;
        uint32_t c = cnt;
        if (my_off >= p.n_symbols) c = 0;
        else if (my_off + c > p.n_symbols) c = (uint32_t)(p.n_symbols - my_off);
        const uint32_t A = (uint32_t)((reinterpret_cast<uintptr_t>(p.out) + base) & 15u);
        if (c) compact_to_tile(s_stage + t, kTileThreads, c, s_tile, A + excl_t);
        fence_proxy_async_smem();
        __syncthreads();   // (D)

        // ---------------- flush the tile: TMA bulk body + scalar head/tail ----------------
        {
            uint64_t tot = s_total;   // symbols of this tile, clamped to n_symbols below
            uint64_t lim = base >= p.n_symbols ? 0 : p.n_symbols - base;
            if (tot > lim) tot = lim;
            const uint32_t total = (uint32_t)tot;
            uint32_t head = (16u - A) & 15u;
            if (head > total) head = total;
            const uint32_t body = (total - head) & ~15u;
            const uint32_t tail = total - head - body;
            uint8_t *gout = p.out + base;
            if (t == 0) {
                if (body) {
                    tma_bulk_store(gout + head, s_tile + A + head, body);
                    tma_bulk_commit();
                    store_pending = true;
                } else {
                    store_pending = false;
                }
            }
            if (t >= 32 && t < 48) {
                const uint32_t i = t - 32;
                if (i < head) gout[i] = s_tile[A + i];
            } else if (t >= 64 && t < 80) {
                const uint32_t i = t - 64;
                if (i < tail) gout[head + body + i] = s_tile[A + head + body + i];
            }
        }
        par ^= 1;
        // the next iteration reads s_tile_id[par] written before barrier (B)
    }
    if (t == 0 && store_pending) tma_bulk_wait0();
}

// =========================================================================
// K1: synchronisation.  A "chunk" is kChunkSegs consecutive 128-bit segments;
// its result is one gap word: nibble k = bit offset at which the parse enters
// segment 8c+k+1 (nibble 7 = exit of the chunk = entry of chunk c+1).
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
    L.has_long = tb.has_long != 0;
}

// K1a: every thread parses its chunk speculatively from bit 0, then the block
// iterates "take the predecessor's exit, re-parse until the parses merge" until
// nothing changes (intra-block fixed point; ballot/barrier-or detects it).
__global__ void __launch_bounds__(kSyncThreads) hd_sync_chunks_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_exit[kSyncThreads];
    SyncLut L;
    sync_stage_tables(p.table, smem, L);
    __syncthreads();

    const GlobalSegLoader ld{p.units, p.n_units};
    const uint32_t tid = threadIdx.x;
    const uint64_t chunk = (uint64_t)blockIdx.x * kSyncThreads + tid;
    const bool active = chunk < p.n_chunks;
    uint32_t entry = (chunk == 0) ? p.first_bit : 0u;
    uint32_t word = 0;
    bool merged;
    if (active) word = sync_parse_chunk(ld, chunk, entry, L, false, 0u, merged);
    s_exit[tid] = word >> 28;
    for (;;) {
        __syncthreads();
        const uint32_t xin = (tid == 0 || !active) ? entry : s_exit[tid - 1];
        const bool redo = xin != entry;
        __syncthreads();
        if (redo) {
            word = sync_parse_chunk(ld, chunk, xin, L, true, word, merged);
            entry = xin;
            s_exit[tid] = word >> 28;
        }
        if (!__syncthreads_or(redo ? 1 : 0)) break;
    }
    if (active) p.gap_words[chunk] = word;
    if (tid == 0) p.blk_entry[blockIdx.x] = (uint8_t)entry;
}

// K1b: one thread per block seam.  Takes the (current) exit of the previous
// block and re-parses the head of its own block until the parses merge.  If
// the walk reaches the end of the block without merging, the block's exit has
// changed and `changed` is bumped: the host launches another round (fixed
// point; degrades to a sequential chain for codes that never synchronise).
__global__ void __launch_bounds__(kSyncThreads) hd_sync_seams_kernel(const SyncParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    SyncLut L;
    sync_stage_tables(p.table, smem, L);
    __syncthreads();
    const GlobalSegLoader ld{p.units, p.n_units};
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;   // seam before block b
    if (b >= p.n_blocks) return;
    const uint64_t c0 = b * kSyncThreads;
    const uint32_t xin = p.gap_words[c0 - 1] >> 28;
    if (xin == p.blk_entry[b]) return;
    uint64_t c_end = c0 + kSyncThreads;
    if (c_end > p.n_chunks) c_end = p.n_chunks;
    uint32_t entry = xin;
    bool merged = false;
    for (uint64_t c = c0; c < c_end; ++c) {
        uint32_t word = p.gap_words[c];
        word = sync_parse_chunk(ld, c, entry, L, true, word, merged);
        p.gap_words[c] = word;
        if (merged) break;
        entry = word >> 28;
    }
    p.blk_entry[b] = (uint8_t)xin;
    if (!merged) atomicAdd(p.changed, 1u);
}

}  // namespace hd
#endif  // HD_KERNELS_CUH_
