/*
 * llhuff.h -- source-compatible surface of the reference's length-limited
 * Huffman encoder (gpuhd/encoder/include/llhuffman_encoder.h:31-42,
 * llhuffman_encoder_table.h) over the host-side producers of hd_b200.h.
 * Input side only (NOT the hot path): lets the demo CLIs make streams without
 * the reference tree.  Unlike the reference packer the tail of the
 * last codeword is always flushed.
 */
#ifndef LLHUFF_SHIM_H_
#define LLHUFF_SHIM_H_

#include <memory>
#include <unordered_map>
#include <vector>

#include "cuhd.h"

namespace llhuff {

struct LLHuffmanEncoderTable {
    struct LLHuffmanEncoderTableItem {
        UNIT_TYPE codeword;
        size_t length;
    };
    size_t compressed_size;   // units
    std::unordered_map<SYMBOL_TYPE, LLHuffmanEncoderTableItem> dict;
    std::vector<hd_sym_len> pairs;   // canonical order (non-decreasing length)
};

class LLHuffmanEncoder {
   public:
    struct Symbol {
        size_t length;
        size_t count;
        SYMBOL_TYPE symbol;
    };

    static std::shared_ptr<std::vector<Symbol>> get_symbol_lengths(SYMBOL_TYPE *data, size_t size) {
        hd_sym_len pairs[256];
        uint32_t n = 0;
        if (hd_build_code(data, size, MAX_CODEWORD_LENGTH, pairs, &n) != HD_OK) return nullptr;
        size_t hist[256] = {0};
        for (size_t i = 0; i < size; ++i) ++hist[data[i]];
        auto v = std::make_shared<std::vector<Symbol>>();
        for (uint32_t i = 0; i < n; ++i) v->push_back(Symbol{pairs[i].length, hist[pairs[i].symbol], pairs[i].symbol});
        return v;
    }

    static std::shared_ptr<LLHuffmanEncoderTable> get_encoder_table(std::shared_ptr<std::vector<Symbol>> lengths) {
        auto t = std::make_shared<LLHuffmanEncoderTable>();
        // canonical codes in list order: code = (code + 1) << (next_len - len)  (llhuffman_encoder.cc:183-195)
        uint32_t code = 0;
        size_t prev = 0, bits = 0;
        for (size_t i = 0; i < lengths->size(); ++i) {
            const Symbol &s = (*lengths)[i];
            if (i) code = (code + 1) << (s.length - prev);
            prev = s.length;
            t->dict[s.symbol] = {code, s.length};
            t->pairs.push_back(hd_sym_len{s.symbol, (uint8_t)s.length});
            bits += s.length * s.count;
        }
        t->compressed_size = SDIV(SDIV(bits, 8), sizeof(UNIT_TYPE));
        return t;
    }

    static void encode_memory(UNIT_TYPE *out, size_t size_out, SYMBOL_TYPE *in, size_t size_in,
                              std::shared_ptr<LLHuffmanEncoderTable> t) {
        hd_encode(in, size_in, t->pairs.data(), (uint32_t)t->pairs.size(), out, size_out, 0, nullptr, 0);
    }

    static std::shared_ptr<cuhd::CUHDCodetable> get_decoder_table(std::shared_ptr<LLHuffmanEncoderTable> t) {
        auto table = std::make_shared<cuhd::CUHDCodetable>((size_t)1 << MAX_CODEWORD_LENGTH);
        hd_flat_lut_from_lengths(t->pairs.data(), (uint32_t)t->pairs.size(), MAX_CODEWORD_LENGTH,
                                 reinterpret_cast<hd_lut_entry *>(table->get()));
        return table;
    }
};

}  // namespace llhuff
#endif /* LLHUFF_SHIM_H_ */
