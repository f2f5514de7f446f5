// hd_decode.cuh -- K4, the decode-and-write kernel of the B200 Huffman decoder.
//
//   hd_decode_tiles_kernel<SEGS, RCAP, LONG>: gap-array segment decode + single-pass
//   decoupled look-back prescan + decode-and-write in ONE persistent launch.
//   Replaces gpu_dec_l1_l2 (Huffman_coding_Gap_arrays/decoder/src/decoder.cu:454-730)
//   and phases 3+4 of CUHD (gpuhd/src/cuhd_gpu_decoder.cu:329-420 + thrust scan
//   :498-505).  Integer/byte work only: no tensor cores.
//
// One persistent CTA per SM: NW decode warps + one scan warp.  A TILE is NW * 32
// chunks of SEGS consecutive 128-bit segments (one chunk per decode thread); tile
// ids come from an atomic ticket counter so the look-back only ever waits for
// tiles that are running.  Every decode warp owns a private block of shared
// memory (no decode warp ever touches another warp's block, so the decode warps
// never synchronise with each other):
//
//   [ input rows: 4*SEGS+2 rows of 32 words ]   lane l's units in column l
//   [ staging buffer 0: stage_rows rows      ]   lane l's symbols in column l,
//   [ staging buffer 1: stage_rows rows      ]   compacted in place, flushed by TMA
//
//   decode warp, step k:  store the (register-prefetched) units of tile k into its
//                         input columns, decode its chunk into staging[k&1]
//                         (decode_chunk<32>), warp-scan the symbol counts, add the
//                         warp total to the tile aggregate (the last warp to finish
//                         publishes it), bar.arrive FULL[k&1];
//                         prefetch tile k+1; wait READY[(k-1)&1]; compact
//                         staging[(k-1)&1] in place and flush the warp's span with
//                         ONE TMA bulk store (SASS UBLKCP) + <= 15 head/tail bytes.
//   scan warp, step k:    bar.sync FULL[k&1]; take the ticket for step k+2;
//                         exclusive scan of the warp totals; decoupled look-back
//                         over the predecessors' 64-bit flagged status words;
//                         publish the inclusive prefix; arrive READY[k&1].
//
// The look-back of tile k runs while the decode warps decode tile k+1, so its L2
// round trips are never on a decode warp's critical path.  Worst-case staging
// (every codeword of minimum length) is sized per WARP and there are no separate
// output tiles, which is what lets ~27 decode warps share one SM.
#ifndef HD_DECODE_CUH_
#define HD_DECODE_CUH_

#include "hd_kernels.cuh"

namespace hd {

struct TileParams {
    const uint32_t *units;
    uint64_t n_units;            // units that may be READ (includes a right halo when the stream continues)
    const uint32_t *gap_words;   // element i = start offset of segment i+1
    uint64_t n_segments;         // segments to DECODE
    uint32_t first_bit;          // start offset of segment 0 ...
    const uint32_t *entry_word;  // ... unless set: start offset of segment 0 = *entry_word >> 28
    uint64_t *shard_info;        // optional {count, entry_bit | unconverged << 32, exit_bit} for the multi-GPU seam check
    const uint32_t *unconverged; // optional: the self-sync pass's "an exit moved after it was published" counter
    TableDev table;
    uint8_t *out;
    uint64_t n_symbols;
    const uint64_t *out_offset_in;   // optional: decode to out + *out_offset_in (capacity n_symbols counts from `out`)
    uint64_t *out_offset_out;        // optional: receives *out_offset_in + symbols written (chains pipelined pieces)
    uint64_t *tile_status;       // n_tiles words (stride kStatusStride), zeroed before launch
    uint32_t *tile_counter;      // zeroed before launch; [1] = look-back gave up (see lookback_wide)
    uint64_t n_tiles;
    uint64_t *count_out;         // optional: total symbols written
    uint32_t stage_rows;         // rows of one staging buffer (one 32-bit word per lane each)
    uint32_t n_warps;            // decode warps per CTA (blockDim.x == (n_warps + 1) * 32)
    uint32_t k_one, k_sel_len, k_sel_cnt8, k_pair_slope;   // 1, kSelLen, kSelCnt8, kPairSelSlope (see decode_chunk)
};

constexpr uint32_t kWarpRowBytes = 128;   // one row of a warp block: 32 lanes x 4 bytes

// inclusive warp scan: the shuffle's own predicate (source lane in range) guards the add -- no lane comparisons
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t;\n\t"
            "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
            "@p add.u32 %0, %0, t;\n\t"
            "}"
            : "+r"(v)
            : "r"(d));
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
    uint32_t r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}

// waits for a phase of a shared-memory barrier; the hint lets the hardware keep the warp suspended for up to ~10 us per
// attempt instead of returning (and being re-issued) every few hundred cycles
__device__ __forceinline__ void mbar_wait_suspended(unsigned long long *bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(a), "r"(parity), "r"(10000u)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ uint32_t smem_ld_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void tma_bulk_store_s(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}

template <int SEGS, int RCAP, bool LONG>
__global__ void __launch_bounds__(896, 1) hd_decode_tiles_kernel(const TileParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr uint32_t kInRows = 4 * SEGS + 2;
    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31, warp = t >> 5;
    const uint32_t nw = p.n_warps;
    const uint32_t nthreads = (nw + 1u) * 32u;
    const uint32_t lm = p.table.lut_bits;
    const uint32_t lut_bytes = 4u << lm;
    // where this launch's output starts (chained pieces: written by the previous launch on the stream)
    const uint64_t out_off = p.out_offset_in ? *p.out_offset_in : 0ull;
    uint8_t *const out = p.out + out_off;
    const uint64_t n_symbols = p.n_symbols > out_off ? p.n_symbols - out_off : 0ull;

    uint32_t *s_lut = reinterpret_cast<uint32_t *>(smem);
    uint8_t *s_order = smem + lut_bytes;
    LongTables *s_long = reinterpret_cast<LongTables *>(s_order + 256);
    const uint32_t tab_bytes = lut_bytes + (LONG ? 256u + ((uint32_t)(sizeof(LongTables) + 15) & ~15u) : 0u);
    // warp blocks start on a row boundary of the shared window (decode_chunk's address arithmetic needs it)
    const uint32_t smem_abs = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t blocks_abs = (smem_abs + tab_bytes + kWarpRowBytes - 1u) & ~(kWarpRowBytes - 1u);
    const uint32_t stage_bytes = p.stage_rows * kWarpRowBytes;
    const uint32_t warp_bytes = kInRows * kWarpRowBytes + 2u * stage_bytes;

    __shared__ uint32_t s_wtot[2][32];   // warp totals of the tile in buffer b; the scan warp turns them into exclusive offsets
    __shared__ uint32_t s_tileid[3];     // ring: tile of step k in slot k % 3
    __shared__ unsigned long long s_gbase[2];   // global ADDRESS at which the tile in buffer b starts (out + exclusive prefix)
    __shared__ uint32_t s_lim[2];               // symbols of the output capacity left at that address (saturated to 32 bits)
    __shared__ uint32_t s_sum[2], s_arrived[2];   // tile total / number of warps that added theirs
    // s_ready[b] completes its n-th phase once the prefix of step k = b + 2n is in s_gbase[b] / s_lim[b].  Parity waits are safe:
    // the scan warp cannot finish step k+2 (the next phase of the same barrier) before every decode warp has arrived
    // at FULL of step k+1, i.e. has long left its wait for step k.
    __shared__ __align__(8) unsigned long long s_ready[2];
    __shared__ uint32_t s_t1_ready;      // the tile of step 1 is in s_tileid[1]

    // Tickets: every tile id comes from the atomic counter, so tiles are only ever claimed by CTAs that are running.
    // The FIRST ticket is taken here, before the tables are staged, the SECOND by the scan warp after that: by then
    // every CTA has its first tile, so a CTA's first look-back depends only on tiles decoded in the same step.
    if (t == nw * 32u) s_tileid[0] = atomicAdd(p.tile_counter, 1u);
    if (t < 2) { s_sum[t] = 0; s_arrived[t] = 0; mbar_init(&s_ready[t], 1u); }
    if (t == 0) s_t1_ready = 0;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.table.blob);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t i = t; i < lut_bytes / 16; i += nthreads) dst[i] = __ldg(src + i);
        if (LONG) {
            if (t < 64) reinterpret_cast<uint32_t *>(s_order)[t] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_order) + t);
            if (t < sizeof(LongTables) / 4) reinterpret_cast<uint32_t *>(s_long)[t] = __ldg(reinterpret_cast<const uint32_t *>(p.table.blob + p.table.off_long) + t);
        }
    }
    __syncthreads();
    // launched right behind the synchronisation kernel (programmatic dependent launch): everything above overlapped its
    // tail; its gap words are read from here on.  (Returns at once for an ordinary launch.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t bar_threads = nthreads;

    if (warp == nw) {
        // ======================= scan warp =======================
        if (lane == 0) {
            s_tileid[1] = atomicAdd(p.tile_counter, 1u);
            __threadfence_block();
            st_release_cta_u32(&s_t1_ready, 1u);
        }
        __syncwarp();
        for (uint32_t k = 0, k3 = 0;; ++k, k3 = k3 == 2u ? 0u : k3 + 1u) {   // k3 = k % 3
            const uint32_t b = k & 1u;
            const uint64_t tile = s_tileid[k3];
            if (tile >= p.n_tiles) break;
            named_bar_sync(kBarFull + b, bar_threads);
            // ticket of step k+2.  Issued here, but no earlier: a tile must not be claimed long before it is decoded,
            // or every higher tile's look-back waits for it.
            uint32_t ticket = 0;
            if (lane == 0) ticket = atomicAdd(p.tile_counter, 1u);
            const uint32_t wt = lane < nw ? s_wtot[b][lane] : 0u;
            const uint32_t wi = warp_incl_scan_u32(wt);
            if (lane < nw) s_wtot[b][lane] = wi - wt;   // exclusive warp offsets
            const uint64_t agg = __shfl_sync(0xffffffffu, wi, 31);
            // (the tile AGGREGATE was already published by the last decode warp to finish; tile 0 published its PREFIX)
            uint64_t excl = 0;
            if (HD_WHATIF & 1) {
                excl = tile * agg;   // (timing experiment: no look-back)
            } else if (tile > 0) {
                excl = lookback_wide(p.tile_status, tile, lane, p.tile_counter + 1);
                if (lane == 0) red_max_gpu_u64(p.tile_status + tile * kStatusStride, kFlagPrefix | (excl + agg));
            }
            if (lane == 0) {
                s_gbase[b] = (unsigned long long)(uintptr_t)(out + excl);
                const uint64_t left = excl < n_symbols ? n_symbols - excl : 0ull;
                s_lim[b] = left > 0xffffffffull ? 0xffffffffu : (uint32_t)left;
                if (tile + 1 == p.n_tiles) {
                    uint64_t tot = excl + agg;
                    tot = tot < n_symbols ? tot : n_symbols;
                    if (p.count_out) *p.count_out = tot;
                    if (p.out_offset_out) *p.out_offset_out = out_off + tot;
                    if (p.shard_info) {
                        p.shard_info[0] = tot;
                        // (the counter was written by the synchronisation kernel, earlier on this stream; nonzero = the
                        // decode has to be repeated after the K1b rounds: every rank sees that in the exchanged triple)
                        const uint64_t unconv = (p.unconverged && *reinterpret_cast<const volatile uint32_t *>(p.unconverged)) ? 1ull << 32 : 0ull;
                        p.shard_info[1] = (uint64_t)(p.entry_word ? (__ldg(p.entry_word) >> 28) : p.first_bit) | unconv;
                        p.shard_info[2] = gap_nibble(p.gap_words, p.n_segments - 1);   // where the parse leaves the shard
                    }
                }
                s_tileid[k3 == 0u ? 2u : k3 - 1u] = ticket;   // slot (k+2)%3 held step k-1's tile, which nobody reads any more
            }
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
    // constants of the decode loop come in as kernel parameters: opaque to ptxas, so they stay constant-bank operands
    // instead of being re-materialised inside the loop (or, for `one`, folded into selects on the ALU pipe)
    T.one = p.k_one;
    T.sel_len = p.k_sel_len;
    T.sel_cnt8 = p.k_sel_cnt8;
    T.pair_slope = p.k_pair_slope;
    T.len_tab = nullptr;
    T.lt = s_long;
    T.order = s_order;
    const uint32_t blk_abs = blocks_abs + warp * warp_bytes;
    const uint32_t col_abs = blk_abs + lane * 4u;                         // this lane's input column
    const uint32_t *const col = reinterpret_cast<const uint32_t *>(smem + (col_abs - smem_abs));
    const uint32_t stage0_abs = blk_abs + kInRows * kWarpRowBytes;
    uint32_t store_pending = 0;   // bit b: a TMA store may still be reading staging buffer b
    bool prev_valid = false;
    uint32_t prev_cnt = 0, prev_incl = 0;
    // warps whose 32 chunks are all complete, with their look-ahead units, and that do not hold chunk 0
    const uint64_t chunks_seg = p.n_segments / SEGS;
    const uint64_t chunks_unit = p.n_units >= 2 ? (p.n_units - 2) / (4u * SEGS) : 0;
    const uint64_t fast_chunks = chunks_seg < chunks_unit ? chunks_seg : chunks_unit;

    // Inputs of the NEXT tile are prefetched into registers while the previous tile is compacted and flushed, so
    // that the DRAM latency of the unit loads is off the critical path of a step.
    uint4 pf_w[SEGS];
    uint2 pf_la = make_uint2(0u, 0u);
    // raw gap words + nibble shifts; the nibbles are extracted when the step starts so that the prefetch itself never
    // waits for a load.  pf_bits == 0: this thread has no chunk in the tile.
    uint32_t pf_gw0 = 0, pf_gw1 = 0, pf_sh = 0, pf_bits = 0;
    auto prefetch = [&](uint64_t tile) {
        pf_bits = 0;
        if (tile >= p.n_tiles) return;
        const uint64_t cw = (tile * nw + warp) * 32u;   // first chunk of this warp
        const uint64_t g0 = (cw + lane) * SEGS;
        if (cw != 0 && cw + 32u <= fast_chunks) {       // interior warp (uniform test)
            const uint32_t *u = p.units + g0 * 4;
#pragma unroll
            for (int q = 0; q < SEGS; ++q) pf_w[q] = __ldg(reinterpret_cast<const uint4 *>(u) + q);
            pf_la = __ldg(reinterpret_cast<const uint2 *>(u + 4 * SEGS));
            const uint64_t j1 = g0 + SEGS - 1;
            pf_gw0 = __ldg(p.gap_words + ((g0 - 1) >> 3));
            pf_gw1 = __ldg(p.gap_words + (j1 >> 3));
            pf_sh = (4u * ((uint32_t)(g0 - 1) & 7u)) | (4u * ((uint32_t)j1 & 7u)) << 8;
            pf_bits = SEGS * kSegmentBits;
            return;
        }
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

    for (uint32_t k = 0, k3 = 0;; ++k, k3 = k3 == 2u ? 0u : k3 + 1u) {   // k3 = k % 3
        const uint32_t b = k & 1u;
        const uint64_t tile = s_tileid[k3];
        const bool valid = tile < p.n_tiles;
        uint32_t cnt = 0, incl = 0;
        if (valid) {
            // this warp's TMA store of two steps ago read the staging buffer that is written again now
            if (store_pending >> b & 1u) {
                tma_bulk_wait_read0();   // (only lane 0 has groups to wait for)
                store_pending = 0;
                __syncwarp();
            }
            // ---------------- decode this thread's chunk (units were prefetched) ----------------
            if (pf_bits) {
                const uint32_t pf_entry = (pf_gw0 >> (pf_sh & 31u)) & 15u;
                const uint32_t pf_stop = pf_bits + ((pf_gw1 >> (pf_sh >> 8)) & 15u);
#pragma unroll
                for (int q = 0; q < SEGS; ++q) {
                    smem_st_u32(col_abs + (4 * q + 0) * kWarpRowBytes, pf_w[q].x, nullptr);
                    smem_st_u32(col_abs + (4 * q + 1) * kWarpRowBytes, pf_w[q].y, nullptr);
                    smem_st_u32(col_abs + (4 * q + 2) * kWarpRowBytes, pf_w[q].z, nullptr);
                    smem_st_u32(col_abs + (4 * q + 3) * kWarpRowBytes, pf_w[q].w, nullptr);
                }
                smem_st_u32(col_abs + (4 * SEGS + 0) * kWarpRowBytes, pf_la.x, nullptr);
                smem_st_u32(col_abs + (4 * SEGS + 1) * kWarpRowBytes, pf_la.y, nullptr);
                if (HD_WHATIF & 4) cnt = 43u;   // (timing experiment: no decode loop)
                else cnt = decode_chunk<32, LONG>(col, pf_entry, pf_stop, T,
                                                  stage0_abs + b * stage_bytes + ((HD_WHATIF & 64) ? ((lane * 43u) & ~3u) : lane * 4u), nullptr);
            }
            // ---------------- warp-wide inclusive scan of the counts ----------------
            incl = warp_incl_scan_u32(cnt);
            if (lane == 31) {
                s_wtot[b][warp] = incl;
                // the last warp to finish publishes the tile aggregate right away (tile 0: it is its own prefix)
                atomicAdd(&s_sum[b], incl);
                __threadfence_block();
                if (atomicAdd(&s_arrived[b], 1u) == nw - 1u) {
                    const uint32_t total = atomicExch(&s_sum[b], 0u);
                    s_arrived[b] = 0;
                    red_max_gpu_u64(p.tile_status + tile * kStatusStride, (tile == 0 ? kFlagPrefix : kFlagAggregate) | (uint64_t)total);
                }
                __threadfence_block();
            }
            __syncwarp();
            named_bar_arrive(kBarFull + b, bar_threads);
        }
        if (k == 0) {
            while (ld_acquire_cta_u32(&s_t1_ready) == 0) __nanosleep(100);   // the scan warp's second ticket
        }
        // the scan warp's prefix of tile k-1 (computed while this warp decoded tile k); its arrival also publishes the
        // ticket of step k+1 (k == 0: s_t1_ready above)
        if (prev_valid) mbar_wait_suspended(&s_ready[b ^ 1u], ((k - 1u) >> 1) & 1u);
        if (valid) prefetch(s_tileid[k3 == 2u ? 0u : k3 + 1u]);
        if (prev_valid) {
            // ---------------- finish tile k-1: compact this warp's symbols in place and flush them ----------------
            const uint32_t pb = b ^ 1u;
            const uint32_t woff = s_wtot[pb][warp];                            // this warp's offset in the tile
            const uint32_t lim = s_lim[pb];                                    // output capacity left at the tile start
            uint8_t *const gp = reinterpret_cast<uint8_t *>((uintptr_t)s_gbase[pb]) + woff;   // the warp's destination
            const uint32_t wtot = __shfl_sync(0xffffffffu, prev_incl, 31);     // this warp's symbols
            const uint32_t excl_l = prev_incl - prev_cnt;
            // clamp to n_symbols (pad bits decode to garbage: the reference clamps with out_pos < next_out_pos,
            // cuhd_gpu_decoder.cu:101-105)
            const uint32_t room = lim > woff ? lim - woff : 0u;
            const uint32_t total = wtot < room ? wtot : room;
            uint32_t c = prev_cnt;
            if (excl_l + c > total) c = excl_l >= total ? 0u : total - excl_l;
            const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gp) & 15u);
            const uint32_t sblk = stage0_abs + pb * stage_bytes;
            const uint32_t rows_max = warp_max_u32((prev_cnt + 3u) >> 2);
            uint32_t w[RCAP + 2];
#pragma unroll
            for (int i = 0; i < RCAP + 2; ++i) w[i] = 0;
            const uint32_t D = sblk + phase + excl_l;
            if (!(HD_WHATIF & 2)) {   // (timing experiment: no compaction)
                column_read<RCAP>(sblk + lane * 4u, rows_max, w, nullptr);
                __syncwarp();
                column_write<RCAP>(w, c, D, rows_max, nullptr);
                __syncwarp();
                column_write_head(w[0], c, D, nullptr);
            }
            fence_proxy_async_smem();
            __syncwarp();
            // flush: the warp's span is contiguous in the block and in the output, and both have the same 16-byte
            // phase -> one TMA bulk store for the aligned body, lanes 0-15 / 16-31 for head / tail bytes
            const uint32_t sp = sblk + phase;
            uint32_t head = (16u - phase) & 15u;
            if (head > total) head = total;
            const uint32_t body = (total - head) & ~15u;
            const uint32_t tail = total - head - body;
            if (body && !(HD_WHATIF & 8)) {   // (warp-uniform; timing experiment 8: no flush)
                if (lane == 0) {
                    tma_bulk_store_s(gp + head, sp + head, body);
                    tma_bulk_commit();
                }
                store_pending |= 1u << pb;
            }
            const uint32_t hb = lane < 16 ? lane : head + body + lane - 16u;   // byte of the span this lane may copy
            if (!(HD_WHATIF & 8) && (lane < 16 ? lane < head : lane - 16u < tail)) gp[hb] = (uint8_t)smem_ld_u8(sp + hb);
        }
        if (!valid) break;
        prev_valid = true;
        prev_cnt = cnt;
        prev_incl = incl;
    }
    tma_bulk_wait0();
}

}  // namespace hd
#endif  // HD_DECODE_CUH_
