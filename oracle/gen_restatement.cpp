/*
 * gen_restatement.cpp -- seeded restatement of the reference's synthetic data
 * generator (TEST INFRASTRUCTURE ONLY, see hd_oracle.h).
 *
 * Follows /root/reference/generate.cpp:32-47 call for call (mt19937_64,
 * generate_canonical<double,10>, uniform_int_distribution<int>(0,3)/(0,255));
 * the only change is that the engine is seeded from the caller instead of
 * std::random_device (generate.cpp:32 -- reference defect D9 makes the
 * original non-reproducible).
 */
#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <random>

extern "C" void oracle_generate(uint8_t *out, uint64_t n, double redundancy,
                                uint64_t seed) {
    redundancy = std::max(0.0, std::min(1.0, redundancy));
    std::mt19937_64 rng(seed);
    std::uniform_int_distribution<int> dist_full(0, 255);
    std::uniform_int_distribution<int> dist_small(0, 3);
    for (uint64_t i = 0; i < n; ++i) {
        double r = std::generate_canonical<double, 10>(rng);
        if (r < redundancy)
            out[i] = static_cast<uint8_t>('A' + dist_small(rng));
        else
            out[i] = static_cast<uint8_t>(dist_full(rng));
    }
}
