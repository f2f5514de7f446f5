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

// E0: length-limited code construction ON THE DEVICE from the 256-bin histogram (replaces the host-side boundary
// package-merge of Huffman_coding_Gap_arrays/encoder/src/package_merge.cpp:107-182 and the sort of
// encoder/src/huff.cpp:189-194): one CTA; the symbols are rank-sorted in parallel, the package-merge lists (<= 2n items
// per level, 16 levels) are built level by level by one thread -- 256 symbols are a few thousand steps -- and the
// {symbol, length} list comes out in the .huff header order (non-decreasing length, most frequent first).  It also
// writes the encoder's code table {code | len << 24} so that nothing of the code has to visit the host.  Same tie
// rules as hd_build_code_from_histogram (csrc/hd_host.cpp): the results are identical.
constexpr uint32_t kCodeMaxItems = 512;   // a package-merge list holds < 2 * 256 items
struct CodeScratch {                      // global scratch of the kernel (hd_workspace_bytes_code)
    unsigned long long w[kMaxCodeLen][kCodeMaxItems];
    uint8_t pkg[kMaxCodeLen][kCodeMaxItems];
    uint32_t size[kMaxCodeLen];
};

__global__ void __launch_bounds__(256) hd_build_code_kernel(const unsigned long long *hist, uint32_t max_len, hd_sym_len *pairs_out,
                                                            uint32_t *n_pairs_out, uint32_t *code_tab, CodeScratch *sc) {
    __shared__ unsigned long long s_w[256];    // weights of the present symbols, ascending
    __shared__ uint8_t s_sym[256];             // ... and their symbols
    __shared__ uint32_t s_len[256];
    __shared__ uint32_t s_n;
    const uint32_t t = threadIdx.x;
    if (t == 0) s_n = 0;
    __syncthreads();
    // rank sort by (count ascending, symbol descending): rank = number of present symbols that sort before this one
    const unsigned long long mine = hist[t];
    if (mine) {
        uint32_t rank = 0;
        for (uint32_t o = 0; o < 256; ++o) {
            const unsigned long long h = hist[o];
            if (h && (h < mine || (h == mine && o > t))) ++rank;
        }
        s_w[rank] = mine;
        s_sym[rank] = (uint8_t)t;
        atomicAdd(&s_n, 1u);
    }
    if (t < 256) code_tab[t] = 0;
    __syncthreads();
    if (t != 0) return;
    const uint32_t n = s_n;
    if (n == 0 || max_len == 0 || max_len > kMaxCodeLen || (1u << max_len) < n) { *n_pairs_out = 0; return; }
    for (uint32_t i = 0; i < n; ++i) s_len[i] = 0;
    if (n == 1) {
        s_len[0] = 1;
    } else {
        for (uint32_t l = 0; l < max_len; ++l) {
            // packages of the previous level, merged with the items (items first on ties)
            const uint32_t n_pk = l ? sc->size[l - 1] / 2 : 0;
            uint32_t a = 0, b = 0, k = 0;
            while (a < n || b < n_pk) {
                const unsigned long long pw = b < n_pk ? sc->w[l - 1][2 * b] + sc->w[l - 1][2 * b + 1] : 0ull;
                if (b >= n_pk || (a < n && s_w[a] <= pw)) { sc->w[l][k] = s_w[a++]; sc->pkg[l][k] = 0; }
                else { sc->w[l][k] = pw; sc->pkg[l][k] = 1; ++b; }
                ++k;
            }
            sc->size[l] = k;
        }
        uint32_t take = 2 * n - 2;
        for (int l = (int)max_len - 1; l >= 0 && take > 0; --l) {
            if (take > sc->size[l]) take = sc->size[l];
            uint32_t items = 0, pkgs = 0;
            for (uint32_t j = 0; j < take; ++j) (sc->pkg[l][j] ? pkgs : items)++;
            for (uint32_t j = 0; j < items; ++j) s_len[j]++;
            take = 2 * pkgs;
        }
    }
    // header order: most frequent first, then a stable sort by length = counting sort over the lengths
    uint32_t k = 0;
    for (uint32_t len = 1; len <= kMaxCodeLen; ++len)
        for (uint32_t i = n; i-- > 0;)
            if (s_len[i] == len) { pairs_out[k].symbol = s_sym[i]; pairs_out[k].length = (uint8_t)len; ++k; }
    *n_pairs_out = k;
    // canonical codes in list order (code = (code + 1) << (next_len - len), decoder/src/get_table.cpp:17-20)
    uint32_t code = 0;
    for (uint32_t i = 0; i < k; ++i) {
        const uint32_t len = pairs_out[i].length;
        code_tab[pairs_out[i].symbol] = code | (len << 24);
        const uint32_t next = i + 1 < k ? pairs_out[i + 1].length : len;
        code = (code + 1) << (next - len);
    }
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
