/*
 * sequential_ref.cpp -- extern "C" door onto the UNMODIFIED reference CPU
 * codec HuffmanSequential (sequential.cpp:14-232), compiled where it lies
 * (the reference main() is renamed out of the way).  TEST INFRASTRUCTURE ONLY.
 * Mirrors sequential.cpp:243-266 (encode, decode, compare).
 */
#define main ref_sequential_cli_main
#include "sequential.cpp"
#undef main

#include <cstring>

extern "C" {

/* round trip through the reference codec. Returns 1 if decode(encode(x))==x.
 * decoded_out (may be NULL) receives n bytes. Times in seconds. */
int ref_sequential_roundtrip(const uint8_t *data, uint64_t n,
                             uint8_t *decoded_out, uint64_t *compressed_bytes,
                             double *encode_s, double *decode_s) {
    HuffmanSequential h;
    std::vector<uint8_t> in(data, data + n);
    auto t0 = std::chrono::high_resolution_clock::now();
    auto comp = h.encode(in);
    auto t1 = std::chrono::high_resolution_clock::now();
    auto dec = h.decode(comp);
    auto t2 = std::chrono::high_resolution_clock::now();
    if (compressed_bytes) *compressed_bytes = comp.size();
    if (encode_s) *encode_s = std::chrono::duration<double>(t1 - t0).count();
    if (decode_s) *decode_s = std::chrono::duration<double>(t2 - t1).count();
    bool ok = dec.size() == in.size() &&
              std::equal(in.begin(), in.end(), dec.begin());
    if (decoded_out && dec.size() == n) std::memcpy(decoded_out, dec.data(), n);
    return ok ? 1 : 0;
}

}
