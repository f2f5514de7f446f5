// decoder.cc -- `./bin/decoder <input.huff> <decoded_file>`
// CLI shape and output lines of the reference gap-array decoder (Huffman_coding_Gap_arrays/decoder/src/huff.cpp:
// 31-129, decoder.cu:803-808), written fresh over hd_b200.h.  Differences on purpose: the second argument
// is honoured, the kernel runs once instead of 200 times, nothing is written past the original size.
#include <cuda_runtime.h>
#include <hd_b200.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

static void fatal(const char *msg) {
    std::fprintf(stderr, "%s\n", msg);
    std::exit(EXIT_FAILURE);
}

int main(int argc, char **argv) {
    if (argc != 3) {
        std::printf("USAGE: %s <input.huff> <decoded_file>\n", argv[0]);
        return EXIT_FAILURE;
    }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) fatal("cannot open input file");
    std::fseek(f, 0, SEEK_END);
    const long fsize = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    void *image = nullptr;
    if (cudaMallocHost(&image, (size_t)fsize + 16) != cudaSuccess) fatal("cudaMallocHost failed");
    if (std::fread(image, 1, (size_t)fsize, f) != (size_t)fsize) fatal("short read");
    std::fclose(f);

    hd_huff_view v;
    if (hd_huff_parse(image, (size_t)fsize, &v) != HD_OK) fatal("malformed .huff file");
    std::printf("Input file: %s\n", argv[1]);
    std::printf("Original size: %u bytes\n", v.original_size);
    std::printf("Compressed size: %u bytes\n", v.n_units * 4);

    void *out = nullptr;
    if (cudaMallocHost(&out, (size_t)v.original_size + 16) != cudaSuccess) fatal("cudaMallocHost failed");
    float ms[3] = {0, 0, 0};
    size_t out_size = 0;
    const auto t0 = std::chrono::high_resolution_clock::now();
    const hd_status st = hd_huff_decode_host(image, (size_t)fsize, static_cast<uint8_t *>(out),
                                             (size_t)v.original_size + 16, &out_size, 0, ms);
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (st != HD_OK) {
        std::fprintf(stderr, "decode failed: %s (%s)\n", hd_status_string(st), hd_last_cuda_error());
        return EXIT_FAILURE;
    }
    std::printf("HtoD,%f, dec,%f, DtoH,%f,SEGMENTSIZE,128,THREAD_NUM,512,LOCAL_SEGMENT_NUM,2\n", ms[0], ms[1], ms[2]);
    const double total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    std::printf("Decode time: %.3f ms\n", total_ms);
    std::printf("Throughput: %.2f MB/s\n", (double)out_size / (1024.0 * 1024.0) / (total_ms / 1000.0));

    FILE *o = std::fopen(argv[2], "wb");
    if (!o) fatal("cannot open output file");
    if (std::fwrite(out, 1, out_size, o) != out_size) fatal("short write");
    std::fclose(o);
    cudaFreeHost(out);
    cudaFreeHost(image);
    return 0;
}
