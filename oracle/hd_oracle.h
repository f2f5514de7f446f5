/*
 * hd_oracle.h -- CPU restatement of the reference's Huffman decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * cse375-finalproj-huffman-decoding_b200/, the C-ABI library, the CLIs) may
 * include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or the timed CPU baseline.
 *
 * Parity status: PINNED.  Checked (tests/test_oracle_pinned.py) against
 *   (1) streams + tables produced by the UNMODIFIED reference encoder
 *       (gpuhd/encoder/src/llhuffman_encoder.cc, built into oracle/_ref by
 *       oracle/Makefile) -- committed as tests/golden/llhuff_*.bin, and
 *   (2) the reference sequential.cpp round trip on the same bytes.
 *   (3) .huff files written by the reference GPU encoder rebuilt for sm_100
 *       on the B200 box (tests/golden/gap_*.huff; script oracle/make_golden_gpu.sh)
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef HD_ORACLE_H_
#define HD_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* gpuhd/include/cuhd_codetable.h:20-23 : {num_bits, symbol} */
typedef struct {
    uint8_t num_bits;
    uint8_t symbol;
} oracle_lut_entry;

/* Canonical code assignment over a (symbol,length) list given in
 * non-decreasing length order: code=0 for the first entry,
 * code=(code+1)<<(next_len-len) afterwards.
 *   gpuhd/encoder/src/llhuffman_encoder.cc:183-195
 *   Huffman_coding_Gap_arrays/decoder/src/get_table.cpp:17-20
 *   Huffman_coding_Gap_arrays/encoder/src/package_merge.cpp:168-181
 * codes[s]/lens[s] are indexed by symbol value (256 entries, len 0 = absent).
 * Returns 0, or -1 if lengths are not non-decreasing / out of range. */
int oracle_canonical_codes(const uint8_t *syms, const uint8_t *lengths, int n,
                           uint32_t codes[256], uint8_t lens[256]);

/* Flat decode LUT, 2^lut_bits entries, index = next lut_bits stream bits
 * (MSB first), every code replicated 2^(lut_bits-len) times.
 *   gpuhd/encoder/src/llhuffman_encoder.cc:240-262
 * Entries not covered by any code keep num_bits=0. */
void oracle_build_flat_lut(const uint32_t codes[256], const uint8_t lens[256],
                           uint32_t lut_bits, oracle_lut_entry *lut);

/* MSB-first packing into 32-bit units, first code bit = bit 31 of unit 0.
 *   gpuhd/encoder/src/llhuffman_encoder.cc:200-238  (flush_tail=0 reproduces
 *     reference defect D1: a last codeword that straddles a unit boundary
 *     loses its tail bits)
 *   Huffman_coding_Gap_arrays/encoder/src/encoder.cu:281-347 (flush_tail=1)
 * gap_nibbles (may be NULL) receives one byte per 128-bit segment:
 *   for a codeword starting at bit p and ending at bit e=p+len with
 *   e/128 != p/128:  gap[p/128] = e & 15        (encoder.cu:307-312)
 * Returns the total number of code bits. units_out must hold
 * ceil(bits/32) units (zero-initialised by this function up to cap_units). */
uint64_t oracle_pack(const uint8_t *in, uint64_t n, const uint32_t codes[256],
                     const uint8_t lens[256], uint32_t *units_out,
                     uint64_t cap_units, int flush_tail, uint8_t *gap_nibbles,
                     uint64_t cap_segments);

/* Pack 8 4-bit gaps per uint32: element i -> word i/8, bits [4(i%8),4(i%8)+3]
 *   Huffman_coding_Gap_arrays/encoder/src/encoder.cu:358-379 */
void oracle_pack_gap_words(const uint8_t *gap_nibbles, uint64_t n_segments,
                           uint32_t *words_out);

/* Sequential flat-LUT decode: look up the top lut_bits of the window, emit
 * the symbol, advance num_bits.
 *   gpuhd/src/cuhd_gpu_decoder.cu:93-117
 * Reads bits past n_units as zero (the reference appends one zero unit,
 * gpuhd/src/cuhd_input_buffer.cc:17-31).  Returns symbols written
 * (== n_symbols unless a zero-length entry is hit). */
uint64_t oracle_decode_flat(const uint32_t *units, uint64_t n_units,
                            const oracle_lut_entry *lut, uint32_t lut_bits,
                            uint8_t *out, uint64_t n_symbols);

/* Per-subsequence synchronisation points of the TRUE parse: for every
 * subsequence s of `subseq_bits` bits, start_off[s] = offset (in bits, from the
 * subsequence start) of the first codeword that starts at or after the
 * subsequence start; counts[s] = number of codewords starting inside s.
 * This is what phases 1+2 of the reference converge to
 *   gpuhd/src/cuhd_gpu_decoder.cu:199-224 (sync point = last codeword start).
 * Returns number of subsequences written. */
uint64_t oracle_sync_points(const uint32_t *units, uint64_t n_units,
                            const oracle_lut_entry *lut, uint32_t lut_bits,
                            uint64_t n_symbols, uint32_t subseq_bits,
                            uint8_t *start_off, uint32_t *counts,
                            uint64_t cap_subseq);

/* Gap-array decode: every 128-bit segment k is decoded independently starting
 * at bit 128k+gap[k-1] (segment 0 at bit 0) until the bit position reaches
 * 128(k+1); results are concatenated.  Output is clamped to n_symbols.
 *   Huffman_coding_Gap_arrays/decoder/src/decoder.cu:505-507 (start bit)
 *   Huffman_coding_Gap_arrays/decoder/src/decoder.cu:529-569, 655-728 (loops)
 * Returns symbols written. */
uint64_t oracle_decode_gap(const uint32_t *units, uint64_t n_units,
                           const uint32_t *gap_words, uint64_t n_segments,
                           const oracle_lut_entry *lut, uint32_t lut_bits,
                           uint8_t *out, uint64_t n_symbols);

/* .huff container (little endian, no magic):
 *   u64 symbol_count | symbol_count x {u8 symbol,u8 length} | u32 original
 *   bytes | u32 units | u32 gap elements | gap words | units
 *   writer: Huffman_coding_Gap_arrays/encoder/src/huff.cpp:186-202
 *   reader: Huffman_coding_Gap_arrays/decoder/src/huff.cpp:36-100 */
typedef struct {
    uint64_t symbol_count;
    const uint8_t *sym_len_pairs; /* symbol_count x {symbol,length} */
    uint32_t original_size;
    uint32_t n_units;
    uint32_t n_segments;
    const uint32_t *gap_words; /* ceil(n_segments/8) */
    const uint32_t *units;     /* n_units */
} oracle_huff_view;

/* Returns 0 on success, negative on malformed/truncated input. */
int oracle_huff_parse(const void *file, size_t size, oracle_huff_view *v);

/* Size needed for / write of a .huff image. Returns bytes written. */
size_t oracle_huff_size(uint64_t symbol_count, uint32_t n_units,
                        uint32_t n_segments);
size_t oracle_huff_write(void *dst, const uint8_t *syms, const uint8_t *lengths,
                         uint64_t symbol_count, uint32_t original_size,
                         const uint32_t *units, uint32_t n_units,
                         const uint32_t *gap_words, uint32_t n_segments);

/* Whole-file helper: parse, build canonical table (max length 16), decode via
 * the gap rule.  out must hold original_size bytes.  Returns symbols written
 * or negative on parse error. */
int64_t oracle_huff_decode(const void *file, size_t size, uint8_t *out,
                           uint64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif /* HD_ORACLE_H_ */
