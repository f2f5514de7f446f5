// demo.cc -- `./bin/demo <compute device index> <size of input in megabytes>`
// Same CLI shape and output lines as the reference's gpuhd demo (gpuhd/src/demo.cc:29-42, 61, 84, 97, 170,
// 236-248), written fresh over the cuhd.h / llhuff.h shims: read ../data100_100.bin if present (else generate
// generate.cpp-distribution data at redundancy 0.5 and save it), encode on the host, decode on the GPU through the
// B200 decoder, time, compare.  Exit code 0 either way, "mismatch" printed on failure (reference behaviour).
#include <cuhd.h>
#include <llhuff.h>

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#define INPUT_FILENAME "../data100_100.bin"
#define SUBSEQ_SIZE 4
#define NUM_THREADS 128

int main(int argc, char **argv) {
    const char *usage = "USAGE: demo <compute device index> <size of input in megabytes>";
    if (argc < 3) { std::cout << usage << std::endl; return 1; }
    const std::int64_t compute_device_id = atoi(argv[1]);
    const std::int64_t size_mb = atoi(argv[2]);
    if (compute_device_id < 0 || size_mb < 1) { std::cout << usage << std::endl; return 1; }
    const size_t size = (size_t)size_mb * 1024 * 1024;

    std::vector<std::pair<std::string, size_t>> timings;
    std::vector<SYMBOL_TYPE> buffer(size);
    {
        std::ifstream in(INPUT_FILENAME, std::ios::binary);
        bool loaded = false;
        if (in) {
            in.read(reinterpret_cast<char *>(buffer.data()), (std::streamsize)size);
            loaded = (size_t)in.gcount() == size;
        }
        if (loaded) {
            std::cout << "Input source: Read " << size_mb << " MB from file: " << INPUT_FILENAME << std::endl;
        } else {
            TIMER_START(timings, "generating random data")
                hd_generate(buffer.data(), size, 0.5, 1, 0);
            TIMER_STOP
            std::ofstream out(INPUT_FILENAME, std::ios::binary);
            if (out) {
                out.write(reinterpret_cast<const char *>(buffer.data()), (std::streamsize)size);
                std::cout << "Data saved to '" << INPUT_FILENAME << "'." << std::endl;
            }
        }
    }

    std::shared_ptr<llhuff::LLHuffmanEncoderTable> enc_table;
    std::shared_ptr<cuhd::CUHDCodetable> dec_table;
    std::unique_ptr<UNIT_TYPE[]> compressed;
    TIMER_START(timings, "Compression time")
        auto lengths = llhuff::LLHuffmanEncoder::get_symbol_lengths(buffer.data(), size);
        enc_table = llhuff::LLHuffmanEncoder::get_encoder_table(lengths);
        dec_table = llhuff::LLHuffmanEncoder::get_decoder_table(enc_table);
        compressed = std::make_unique<UNIT_TYPE[]>(enc_table->compressed_size);
        llhuff::LLHuffmanEncoder::encode_memory(compressed.get(), enc_table->compressed_size, buffer.data(), size,
                                                enc_table);
    TIMER_STOP

    cudaSetDevice((int)compute_device_id);
    CUERR

    auto in_buf = std::make_shared<cuhd::CUHDInputBuffer>(reinterpret_cast<std::uint8_t *>(compressed.get()),
                                                          enc_table->compressed_size * sizeof(UNIT_TYPE));
    auto out_buf = std::make_shared<cuhd::CUHDOutputBuffer>(size);
    auto gpu_in = std::make_shared<cuhd::CUHDGPUInputBuffer>(in_buf);
    auto gpu_out = std::make_shared<cuhd::CUHDGPUOutputBuffer>(out_buf);
    auto gpu_table = std::make_shared<cuhd::CUHDGPUCodetable>(dec_table);
    auto gpu_aux = std::make_shared<cuhd::CUHDGPUDecoderMemory>(in_buf->get_compressed_size_units(), SUBSEQ_SIZE,
                                                                NUM_THREADS);
    // HD_DEMO_VERBOSE=1: where the "Decompression time" goes (stderr; the reference times the same region as a whole)
    const bool verbose = std::getenv("HD_DEMO_VERBOSE") != nullptr;
    auto lap = [&, t = std::chrono::steady_clock::now()](const char *what) mutable {
        if (!verbose) return;
        const auto n = std::chrono::steady_clock::now();
        std::cerr << "  " << what << ": " << std::chrono::duration<double, std::milli>(n - t).count() << " ms" << std::endl;
        t = n;
    };
    lap("(start)");
    TIMER_START(timings, "Decompression time")
        gpu_in->allocate();
        gpu_out->allocate();
        gpu_table->allocate();
        gpu_aux->allocate();
        lap("4 x allocate (cudaMalloc)");
        gpu_table->cpy_host_to_device();
        lap("table upload");
        gpu_in->cpy_host_to_device();
        lap("stream H2D");
        cuhd::CUHDGPUDecoder::decode(gpu_in, in_buf->get_compressed_size_units(), gpu_out,
                                     out_buf->get_uncompressed_size(), gpu_table, gpu_aux, MAX_CODEWORD_LENGTH,
                                     SUBSEQ_SIZE, NUM_THREADS);
        lap("decode");
        gpu_out->cpy_device_to_host();
        lap("output D2H");
    TIMER_STOP

    for (auto &t : timings) std::cout << t.first << ".. " << t.second << "µs" << std::endl;
    const size_t n = timings.size();
    std::cout << "Compression throughput (mb/s): " << size / std::max<size_t>(1, timings[n - 2].second) << std::endl;
    std::cout << "Decompression throughput (mb/s): " << size / std::max<size_t>(1, timings[n - 1].second) << std::endl;
    std::cout << std::endl;
    if (!cuhd::CUHDUtil::equals(buffer.data(), out_buf->get_decompressed_data().get(), size))
        std::cout << "mismatch" << std::endl;
    return 0;
}
