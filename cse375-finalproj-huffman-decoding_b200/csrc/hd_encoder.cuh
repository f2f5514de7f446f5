// hd_encoder.cuh -- device-side producer of the .huff payload: byte histogram, parallel
// MSB-first bit packing with a single-pass decoupled look-back over BIT offsets, and the 4-bit gap array.
// Replaces cu_histgram/mergeHistogram (Huffman_coding_Gap_arrays/encoder/src/encoder.cu:33-116), cuencoder (:142-355)
// and cu_get_gaparray (:358-379).  Input side of the hot path, not the hot path itself; written for correctness and a
// few hundred GB/s, not tuned further.  Output is bit-identical to hd_encode() / the oracle packer.
#ifndef HD_ENCODER_CUH_
#define HD_ENCODER_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "hd_kernels.cuh"

namespace hd {

constexpr uint32_t kEncThreads = 256;
constexpr uint32_t kEncSymsPerThread = 16;
constexpr uint32_t kEncTileSyms = kEncThreads * kEncSymsPerThread;   // 4096 symbols per tile

// E1: byte histogram.  One private 256-bin histogram per warp in shared memory (conflicts only inside a warp),
// merged into 256 global 64-bit counters.
__global__ void __launch_bounds__(256) hd_histogram_kernel(const uint8_t *data, uint64_t n, unsigned long long *hist) {
    __shared__ uint32_t s_h[8][256];
    const uint32_t t = threadIdx.x, warp = t >> 5;
    for (uint32_t i = t; i < 8 * 256; i += 256) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t n16 = n / 16;
    const uint4 *d16 = reinterpret_cast<const uint4 *>(data);
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + t; i < n16; i += (uint64_t)gridDim.x * 256) {
        const uint4 v = __ldg(d16 + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(&s_h[warp][w[k] & 0xff], 1u);
            atomicAdd(&s_h[warp][(w[k] >> 8) & 0xff], 1u);
            atomicAdd(&s_h[warp][(w[k] >> 16) & 0xff], 1u);
            atomicAdd(&s_h[warp][w[k] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (uint64_t i = n16 * 16 + t; i < n; i += 256) atomicAdd(&s_h[warp][data[i]], 1u);
    __syncthreads();
    uint32_t sum = 0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sum += s_h[wv][t];
    if (sum) atomicAdd(hist + t, (unsigned long long)sum);
}

struct EncodeParams {
    const uint8_t *data;
    uint64_t n;
    const uint32_t *code_tab;      // device: 256 x {code | len << 24}
    uint32_t *units;               // zeroed before launch
    uint64_t cap_units;
    uint32_t *gap_words;           // zeroed before launch (may be NULL)
    uint64_t *tile_status;         // zeroed; kStatusStride words per tile
    uint32_t *tile_counter;        // zeroed
    uint64_t n_tiles;
    uint64_t *bits_out;            // optional
};

// E2: pack.  Tiles of 4096 symbols taken by ticket; per thread 16 symbols -> their total bit length -> block scan ->
// look-back over tile bit totals -> every thread ORs its bits into the tile's bit buffer in shared memory (first and
// last word with atomics, interior words are its own) -> the tile buffer goes to the unit array (first/last word of
// the tile with a global atomicOr: they are shared with the neighbouring tiles).  A codeword that crosses a 128-bit
// boundary records how far it reaches into the next segment: the gap (encoder.cu:307-312).
__global__ void __launch_bounds__(kEncThreads) hd_encode_kernel(const EncodeParams p) {
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_bits[kEncTileSyms * 16 / 32 + 2];   // tile bit buffer: <= 65536 bits + 31 lead bits
    __shared__ uint32_t s_wtot[kEncThreads / 32];
    __shared__ uint32_t s_tile, s_total;
    __shared__ unsigned long long s_base;
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    s_tab[t] = p.code_tab[t];
    for (;;) {
        __syncthreads();   // previous tile fully written out; s_tab visible
        if (t == 0) s_tile = atomicAdd(p.tile_counter, 1u);
        for (uint32_t i = t; i < sizeof(s_bits) / 4; i += kEncThreads) s_bits[i] = 0;
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= p.n_tiles) return;
        // ---- this thread's 16 symbols and their bit length ----
        const uint64_t s0 = tile * kEncTileSyms + (uint64_t)t * kEncSymsPerThread;
        uint8_t sym[kEncSymsPerThread];
        uint32_t cnt = 0;
        if (s0 + kEncSymsPerThread <= p.n) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.data + s0));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) sym[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            cnt = kEncSymsPerThread;
        } else if (s0 < p.n) {
            cnt = (uint32_t)(p.n - s0);
#pragma unroll
            for (int k = 0; k < 16; ++k) sym[k] = (uint32_t)k < cnt ? p.data[s0 + k] : 0;
        }
        uint32_t len_t = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if ((uint32_t)k < cnt) len_t += s_tab[sym[k]] >> 24;
        // ---- block exclusive scan of the bit lengths ----
        uint32_t incl = len_t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += v;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t wt = lane < kEncThreads / 32 ? s_wtot[lane] : 0u;
            uint32_t wi = wt;
#pragma unroll
            for (int d = 1; d < (int)(kEncThreads / 32); d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += v;
            }
            if (lane < kEncThreads / 32) s_wtot[lane] = wi - wt;
            const uint64_t agg = __shfl_sync(0xffffffffu, wi, kEncThreads / 32 - 1);
            if (lane == 0) red_max_gpu_u64(p.tile_status + tile * kStatusStride, (tile == 0 ? kFlagPrefix : kFlagAggregate) | agg);
            uint64_t excl = 0;
            if (tile > 0) {
                excl = lookback_wide(p.tile_status, tile, lane, p.tile_counter + 1);
                if (lane == 0) red_max_gpu_u64(p.tile_status + tile * kStatusStride, kFlagPrefix | (excl + agg));
            }
            if (lane == 0) {
                s_base = excl;
                if (p.bits_out && tile + 1 == p.n_tiles) *p.bits_out = excl + agg;
            }
        }
        __syncthreads();
        const uint64_t base = s_base;                                  // absolute bit offset of the tile
        const uint32_t lead = (uint32_t)(base & 31u);                  // the tile buffer starts at unit base >> 5
        uint32_t q = lead + s_wtot[warp] + incl - len_t;               // bit index of this thread inside the buffer
        uint64_t abs_bit = base + s_wtot[warp] + incl - len_t;         // absolute bit position (for the gaps)
        // ---- append the codes: 64-bit accumulator, flush whole words ----
        unsigned long long acc = 0;
        uint32_t nacc = q & 31u;                                       // leading zero bits: someone else's
        uint32_t word = q >> 5;
        bool first = true;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if ((uint32_t)k < cnt) {
                const uint32_t e = s_tab[sym[k]];
                const uint32_t len = e >> 24, code = e & 0xffffffu;
                acc = (acc << len) | code;
                nacc += len;
                const uint64_t end_bit = abs_bit + len;
                if (p.gap_words && ((abs_bit ^ end_bit) >> 7)) {       // crosses a 128-bit boundary
                    const uint64_t seg = abs_bit >> 7;
                    const uint32_t gap = (uint32_t)(end_bit & 127u);
                    if (gap) atomicOr(p.gap_words + (seg >> 3), gap << (4u * (uint32_t)(seg & 7)));
                }
                abs_bit = end_bit;
                if (nacc >= 32) {
                    const uint32_t w32 = (uint32_t)(acc >> (nacc - 32));
                    if (first) atomicOr(&s_bits[word], w32); else s_bits[word] = w32;
                    first = false;
                    ++word;
                    nacc -= 32;
                    acc &= (1ull << nacc) - 1ull;
                }
            }
        }
        if (nacc) {
            // (if nothing was flushed the leading zero bits are still in front: the shift places the codes right)
            atomicOr(&s_bits[word], (uint32_t)(acc << (32 - nacc)));
        }
        __syncthreads();
        // ---- tile buffer -> units (total bits of the tile = inclusive offset of the last thread) ----
        if (t == kEncThreads - 1) s_total = s_wtot[warp] + incl;
        __syncthreads();
        const uint32_t n_words = (lead + s_total + 31u) >> 5;
        const uint64_t w0 = base >> 5;
        for (uint32_t i = t; i < n_words; i += kEncThreads) {
            const uint32_t v = s_bits[i];
            if (w0 + i >= p.cap_units) continue;
            if (i == 0 || i + 1 == n_words) { if (v) atomicOr(p.units + w0 + i, v); }
            else p.units[w0 + i] = v;
        }
    }
}

}  // namespace hd
#endif  // HD_ENCODER_CUH_
