// hd_tables.cuh -- K0: decode-table construction ON THE DEVICE from a {symbol,length} list (the header of a .huff
// file, Huffman_coding_Gap_arrays/encoder/src/huff.cpp:189-194).  Replaces the host-side table builders of the
// reference decoders: get_table_info / get_twolevel_table (decoder/src/get_table.cpp:3-139) and
// LLHuffmanEncoder::get_decoder_table (gpuhd/encoder/src/llhuffman_encoder.cc:240-262) -- SURVEY section 8(f) row 1.
// Writes the same blob, byte for byte, as build_table_blob() in hd_host.cpp (the parity test compares them).
#ifndef HD_TABLES_CUH_
#define HD_TABLES_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "hd_internal.h"

namespace hd {

struct TableSummary {      // read back by the host (20 bytes): everything it needs to launch the decode kernels
    uint32_t ok;           // 0 = malformed list
    uint32_t lut_bits;
    uint32_t min_len;
    uint32_t max_len;
    uint32_t has_long;     // codes longer than the LUT, or an incomplete code: the kernels need the resolver
};

constexpr uint32_t kTableBlobMaxBytes = (4u << kMaxLutBits) + (2u << kMaxLutBits) + 256 + 256 + ((sizeof(LongTables) + 15) & ~15u);

// canonical search: the codeword that starts at bit `pos` of the lm-bit window w (right aligned), if it lies inside
__device__ __forceinline__ bool table_first_code(uint32_t w, uint32_t pos, uint32_t lm, const LongTables &lt,
                                                 const uint8_t *order, uint32_t max_len, uint32_t &len, uint32_t &sym) {
    const uint32_t avail = lm - pos;
    const uint32_t rest = w & ((1u << avail) - 1u);
    const uint32_t top = avail < max_len ? avail : max_len;
    for (uint32_t l = 1; l <= top; ++l) {
        if (!lt.count[l]) continue;
        const uint32_t d = (rest >> (avail - l)) - lt.first[l];
        if (d < lt.count[l]) {
            len = l;
            sym = order[lt.offset[l] + d];
            return true;
        }
    }
    return false;
}

__global__ void __launch_bounds__(1024) hd_build_tables_kernel(const hd_sym_len *pairs, uint32_t n_pairs, uint8_t *blob,
                                                               TableSummary *summary) {
    __shared__ LongTables lt;
    __shared__ uint8_t s_len[256], s_order[256];
    __shared__ uint32_t s_ok, s_lm, s_min, s_max, s_long;
    const uint32_t t = threadIdx.x;
    if (t < 256) { s_len[t] = 0; s_order[t] = 0; }
    if (t < 17) { lt.first[t] = 0; lt.count[t] = 0; lt.offset[t] = 0; }
    __syncthreads();
    if (t == 0) {
        // canonical assignment in list order: code = (code + 1) << (next_len - len)  (get_table.cpp:17-20)
        bool ok = n_pairs > 0 && n_pairs <= 256;
        uint64_t code = 0;
        uint32_t min_len = 0, max_len = 0;
        for (uint32_t i = 0; ok && i < n_pairs; ++i) {
            const uint32_t len = pairs[i].length, sym = pairs[i].symbol;
            if (len == 0 || len > kMaxCodeLen || s_len[sym] || (code >> len)) { ok = false; break; }
            s_len[sym] = (uint8_t)len;
            s_order[i] = (uint8_t)sym;
            if (!lt.count[len]) { lt.first[len] = (uint32_t)code; lt.offset[len] = i; }
            lt.count[len] += 1;
            if (!min_len) min_len = len;
            max_len = len;
            const uint32_t next = i + 1 < n_pairs ? pairs[i + 1].length : len;
            if (next < len) { ok = false; break; }
            code = (code + 1) << (next - len);
        }
        uint32_t lm = 2 * max_len < kMaxLutBits ? 2 * max_len : kMaxLutBits;
        if (lm < max_len) lm = max_len < kMaxLutBits ? max_len : kMaxLutBits;
        if (lm < 1) lm = 1;
        s_ok = ok ? 1u : 0u;
        s_lm = lm;
        s_min = min_len;
        s_max = max_len;
        s_long = (max_len > lm || code != (1ull << max_len)) ? 1u : 0u;   // long codes, or an incomplete code (LUT holes)
        summary->ok = s_ok;
        summary->lut_bits = lm;
        summary->min_len = min_len;
        summary->max_len = max_len;
        summary->has_long = s_long;
    }
    __syncthreads();
    if (!s_ok) return;
    const uint32_t lm = s_lm, n = 1u << lm, max_len = s_max;
    uint32_t *dec = reinterpret_cast<uint32_t *>(blob);
    uint16_t *syn = reinterpret_cast<uint16_t *>(blob + n * 4);
    uint8_t *b_len = blob + n * 6, *b_order = b_len + 256;
    for (uint32_t w = t; w < n; w += blockDim.x) {
        uint32_t len0, sym0;
        if (!table_first_code(w, 0, lm, lt, s_order, max_len, len0, sym0)) {
            dec[w] = 0;   // (s_long is set) -> resolve_long finds the long codeword or skips 1 bit
            syn[w] = 0;
            continue;
        }
        uint32_t bits = len0, count = 1, len1, sym1 = 0;
        if (table_first_code(w, bits, lm, lt, s_order, max_len, len1, sym1)) {
            bits += len1;
            count = 2;
        } else {
            sym1 = 0;
        }
        dec[w] = lut_entry(bits, sym0, sym1, count);
        uint32_t jump = len0, l, s;
        while (jump < lm && table_first_code(w, jump, lm, lt, s_order, max_len, l, s)) jump += l;
        syn[w] = (uint16_t)(jump | (len0 << 8));
    }
    if (t < 256) { b_len[t] = s_len[t]; b_order[t] = s_order[t]; }
    if (t < sizeof(LongTables) / 4) reinterpret_cast<uint32_t *>(b_order + 256)[t] = reinterpret_cast<const uint32_t *>(&lt)[t];
}

}  // namespace hd
#endif  // HD_TABLES_CUH_
