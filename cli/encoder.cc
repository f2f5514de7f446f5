// encoder.cc -- `./bin/encoder <input> <output.huff>`
// CLI shape of the reference gap-array encoder (Huffman_coding_Gap_arrays/encoder/src/huff.cpp:31-34, 176-202); a HOST
// producer of the same .huff container (<=16-bit length-limited canonical code, 4-bit gap per 128-bit segment).
// Input side only -- not the hot path (component 7).
#include <hd_b200.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv) {
    if (argc != 3) {
        std::printf("USAGE: %s <input> <output.huff>\n", argv[0]);
        return EXIT_FAILURE;
    }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::fprintf(stderr, "cannot open input file\n"); return EXIT_FAILURE; }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> data((size_t)n);
    if (n && std::fread(data.data(), 1, (size_t)n, f) != (size_t)n) { std::fprintf(stderr, "short read\n"); return EXIT_FAILURE; }
    std::fclose(f);
    if ((unsigned long)n >= 0xffffffffUL) { std::fprintf(stderr, ".huff holds < 4 GiB\n"); return EXIT_FAILURE; }

    const auto t0 = std::chrono::high_resolution_clock::now();
    hd_sym_len pairs[256];
    uint32_t n_pairs = 0;
    uint64_t bits = 0;
    if (hd_build_code(data.data(), (uint64_t)n, 16, pairs, &n_pairs) != HD_OK ||
        hd_encoded_size(data.data(), (uint64_t)n, pairs, n_pairs, &bits) != HD_OK) {
        std::fprintf(stderr, "cannot build a code for this input\n");
        return EXIT_FAILURE;
    }
    const uint32_t n_units = (uint32_t)((bits + 31) / 32), n_seg = (uint32_t)((bits + 127) / 128);
    std::vector<uint32_t> units(n_units ? n_units : 1, 0), gaps((n_seg + 7) / 8 + 1, 0);
    if (hd_encode(data.data(), (uint64_t)n, pairs, n_pairs, units.data(), n_units, 0, gaps.data(), 0) != HD_OK) {
        std::fprintf(stderr, "encode failed\n");
        return EXIT_FAILURE;
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    std::vector<uint8_t> image(hd_huff_image_bytes(n_pairs, n_units, n_seg));
    size_t written = 0;
    if (hd_huff_write(image.data(), image.size(), pairs, n_pairs, (uint32_t)n, units.data(), n_units, gaps.data(), n_seg,
                      &written) != HD_OK) {
        std::fprintf(stderr, "cannot serialise\n");
        return EXIT_FAILURE;
    }
    FILE *o = std::fopen(argv[2], "wb");
    if (!o || std::fwrite(image.data(), 1, written, o) != written) { std::fprintf(stderr, "cannot write output\n"); return EXIT_FAILURE; }
    std::fclose(o);
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    std::printf("Input file: %s\n", argv[1]);
    std::printf("Original size: %u bytes\n", (unsigned)n);
    std::printf("Compressed size: %u bytes\n", n_units * 4);
    std::printf("Total encode time: %.3f ms\n", ms);
    std::printf("Throughput: %.2f MB/s\n", (double)n / (1024.0 * 1024.0) / (ms / 1000.0));
    return 0;
}
