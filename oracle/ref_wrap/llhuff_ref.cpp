/*
 * llhuff_ref.cpp -- thin extern "C" door onto the UNMODIFIED reference encoder
 * (gpuhd/encoder/src/llhuffman_encoder.cc + gpuhd/src/cuhd_codetable.cc),
 * compiled where those sources lie by oracle/Makefile into
 * oracle/_ref/libref_llhuff.so.  TEST INFRASTRUCTURE ONLY.
 *
 * Mirrors the call sequence of gpuhd/src/demo.cc:97-121 and the padding of
 * gpuhd/src/cuhd_input_buffer.cc:17-31.
 */
#include <llhuff.h>
#include <cuhd_input_buffer.h>

#include <cstring>
#include <memory>

extern "C" {

/* returns number of units INCLUDING the appended zero unit (what demo.cc:188
 * passes to decode), or 0 on failure.  lut_out: 2^11 x {num_bits,symbol}. */
uint64_t ref_llhuff_encode(const uint8_t *data, uint64_t n, uint8_t *lut_out,
                           uint32_t *units_out, uint64_t cap_units,
                           uint32_t codes_out[256], uint8_t lens_out[256]) {
    auto lengths = llhuff::LLHuffmanEncoder::get_symbol_lengths(
        const_cast<uint8_t *>(data), n);
    if (!lengths) return 0;
    auto enc = llhuff::LLHuffmanEncoder::get_encoder_table(lengths);
    auto dec = llhuff::LLHuffmanEncoder::get_decoder_table(enc);
    std::unique_ptr<UNIT_TYPE[]> compressed =
        std::make_unique<UNIT_TYPE[]>(enc->compressed_size);
    llhuff::LLHuffmanEncoder::encode_memory(compressed.get(),
        enc->compressed_size, const_cast<uint8_t *>(data), n, enc);
    cuhd::CUHDInputBuffer in(reinterpret_cast<uint8_t *>(compressed.get()),
                             enc->compressed_size * sizeof(UNIT_TYPE));
    uint64_t nu = in.get_compressed_size_units();
    if (nu > cap_units) return 0;
    std::memcpy(units_out, in.get_compressed_data(), nu * sizeof(UNIT_TYPE));
    std::memcpy(lut_out, dec->get(),
                dec->get_size() * sizeof(cuhd::CUHDCodetableItemSingle));
    std::memset(codes_out, 0, 256 * sizeof(uint32_t));
    std::memset(lens_out, 0, 256);
    for (auto &kv : enc->dict) {
        codes_out[kv.first] = kv.second.codeword;
        lens_out[kv.first] = (uint8_t)kv.second.length;
    }
    return nu;
}

uint32_t ref_llhuff_max_codeword_length() { return MAX_CODEWORD_LENGTH; }

}
