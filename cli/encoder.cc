// encoder.cc -- `./bin/encoder [--device] <input> <output.huff>`
// CLI shape of the reference gap-array encoder (Huffman_coding_Gap_arrays/encoder/src/huff.cpp:31-34, 176-202); a
// producer of the same .huff container (<=16-bit length-limited canonical code, 4-bit gap per 128-bit segment).
// Default: the library's host producer.  --device: histogram, length-limited code construction, bit packing and gap
// array all run on the GPU (hd_encode_device_auto, replacing encoder.cu:33-116, 142-379 and the host package-merge of
// package_merge.cpp:107-182) and the device time of those kernels is printed next to the total.
// Input side only -- not the hot path (component 7; section 8(f) row 2).
#include <cuda_runtime.h>
#include <hd_b200.h>

#include <cstring>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define DEV_CHECK(call)                                                                        \
    do {                                                                                       \
        if ((call) != cudaSuccess) { std::fprintf(stderr, "CUDA error in %s\n", #call); return EXIT_FAILURE; } \
    } while (0)
#define HD_CHECK(call)                                                                         \
    do {                                                                                       \
        const hd_status st__ = (call);                                                         \
        if (st__ != HD_OK) { std::fprintf(stderr, "%s: %s (%s)\n", #call, hd_status_string(st__), hd_last_cuda_error()); return EXIT_FAILURE; } \
    } while (0)

// histogram -> length-limited canonical code (kernel) -> pack + gap array: everything on the GPU, one call
static int encode_on_device(const std::vector<uint8_t> &data, hd_sym_len *pairs, uint32_t *n_pairs, std::vector<uint32_t> &units,
                            std::vector<uint32_t> &gaps, uint32_t *n_units, uint32_t *n_seg, float *kernel_ms) {
    const uint64_t n = data.size();
    uint8_t *d_data = nullptr;
    uint32_t *d_units = nullptr, *d_gaps = nullptr;
    uint8_t *d_meta = nullptr;   // [0,512): pairs, [512,516): n_pairs, [520,528): total bits
    void *d_ws = nullptr;
    const uint64_t cap_units = n * 9 / 32 + 16, cap_gaps = (cap_units / 4 + 8) / 8 + 8;
    const size_t ws_bytes = (hd_workspace_bytes_encode(n) + 63) / 64 * 64 + 2048 + hd_workspace_bytes_code();
    DEV_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_data), n + 64));
    DEV_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_meta), 1024));
    DEV_CHECK(cudaMemset(d_meta, 0, 1024));
    DEV_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_units), cap_units * 4));
    DEV_CHECK(cudaMalloc(reinterpret_cast<void **>(&d_gaps), cap_gaps * 4));
    DEV_CHECK(cudaMalloc(&d_ws, ws_bytes));
    DEV_CHECK(cudaMemcpy(d_data, data.data(), n, cudaMemcpyHostToDevice));
    cudaEvent_t ev[2];
    for (auto &e : ev) DEV_CHECK(cudaEventCreate(&e));
    DEV_CHECK(cudaEventRecord(ev[0]));
    HD_CHECK(hd_encode_device_auto(d_data, n, 16, reinterpret_cast<hd_sym_len *>(d_meta), reinterpret_cast<uint32_t *>(d_meta + 512),
                                   d_units, cap_units, d_gaps, cap_gaps, d_ws, ws_bytes, reinterpret_cast<uint64_t *>(d_meta + 520),
                                   nullptr));
    DEV_CHECK(cudaEventRecord(ev[1]));
    DEV_CHECK(cudaEventSynchronize(ev[1]));
    DEV_CHECK(cudaEventElapsedTime(kernel_ms, ev[0], ev[1]));
    uint8_t meta[528];
    DEV_CHECK(cudaMemcpy(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost));
    uint64_t bits = 0;
    std::memcpy(n_pairs, meta + 512, 4);
    std::memcpy(&bits, meta + 520, 8);
    if (*n_pairs == 0 || *n_pairs > 256) { std::fprintf(stderr, "cannot build a code for this input\n"); return 1; }
    std::memcpy(pairs, meta, *n_pairs * sizeof(hd_sym_len));
    *n_units = (uint32_t)((bits + 31) / 32);
    *n_seg = (uint32_t)((bits + 127) / 128);
    units.assign(*n_units ? *n_units : 1, 0);
    gaps.assign((*n_seg + 7) / 8 + 1, 0);
    DEV_CHECK(cudaMemcpy(units.data(), d_units, (size_t)*n_units * 4, cudaMemcpyDeviceToHost));
    DEV_CHECK(cudaMemcpy(gaps.data(), d_gaps, (size_t)((*n_seg + 7) / 8) * 4, cudaMemcpyDeviceToHost));
    for (auto &e : ev) cudaEventDestroy(e);
    cudaFree(d_data); cudaFree(d_meta); cudaFree(d_units); cudaFree(d_gaps); cudaFree(d_ws);
    return 0;
}

int main(int argc, char **argv) {
    const bool on_device = argc == 4 && std::strcmp(argv[1], "--device") == 0;
    if (argc != 3 && !on_device) {
        std::printf("USAGE: %s [--device] <input> <output.huff>\n", argv[0]);
        return EXIT_FAILURE;
    }
    if (on_device) { ++argv; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::fprintf(stderr, "cannot open input file\n"); return EXIT_FAILURE; }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    if (n < 0) { std::fprintf(stderr, "cannot determine the size of the input file\n"); return EXIT_FAILURE; }
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> data((size_t)n);
    if (n && std::fread(data.data(), 1, (size_t)n, f) != (size_t)n) { std::fprintf(stderr, "short read\n"); return EXIT_FAILURE; }
    std::fclose(f);
    if ((unsigned long)n >= 0xffffffffUL) { std::fprintf(stderr, ".huff holds < 4 GiB\n"); return EXIT_FAILURE; }

    if (on_device) cudaFree(nullptr);   // (context creation is not part of the encode)
    const auto t0 = std::chrono::high_resolution_clock::now();
    hd_sym_len pairs[256];
    uint32_t n_pairs = 0, n_units = 0, n_seg = 0;
    std::vector<uint32_t> units, gaps;
    float kernel_ms = 0.f;
    if (on_device) {
        if (n == 0) { std::fprintf(stderr, "empty input\n"); return EXIT_FAILURE; }
        if (encode_on_device(data, pairs, &n_pairs, units, gaps, &n_units, &n_seg, &kernel_ms)) return EXIT_FAILURE;
    } else {
        uint64_t bits = 0;
        if (hd_build_code(data.data(), (uint64_t)n, 16, pairs, &n_pairs) != HD_OK ||
            hd_encoded_size(data.data(), (uint64_t)n, pairs, n_pairs, &bits) != HD_OK) {
            std::fprintf(stderr, "cannot build a code for this input\n");
            return EXIT_FAILURE;
        }
        n_units = (uint32_t)((bits + 31) / 32);
        n_seg = (uint32_t)((bits + 127) / 128);
        units.assign(n_units ? n_units : 1, 0);
        gaps.assign((n_seg + 7) / 8 + 1, 0);
        if (hd_encode(data.data(), (uint64_t)n, pairs, n_pairs, units.data(), n_units, 0, gaps.data(), 0) != HD_OK) {
            std::fprintf(stderr, "encode failed\n");
            return EXIT_FAILURE;
        }
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
    if (on_device)
        std::printf("Device encode kernels (histogram + code + pack + gap array): %.3f ms = %.2f GB/s of input\n", kernel_ms,
                    kernel_ms > 0 ? (double)n / (kernel_ms * 1e-3) / 1e9 : 0.0);
    std::printf("Total encode time: %.3f ms\n", ms);
    std::printf("Throughput: %.2f MB/s\n", (double)n / (1024.0 * 1024.0) / (ms / 1000.0));
    return 0;
}
