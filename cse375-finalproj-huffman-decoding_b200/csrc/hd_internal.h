// hd_internal.h -- shared declarations between the host side (tables, formats,
// encoder) and the CUDA side (kernels + C-ABI entry points).  Not installed.
#ifndef HD_INTERNAL_H_
#define HD_INTERNAL_H_

#include <cstddef>
#include <cstdint>
#include <vector>

#include "hd_b200.h"

namespace hd {

// ---- geometry of the decode-and-write kernel (K4) -------------------------
constexpr uint32_t kSegmentBits = 128;   // .huff SEGMENTSIZE (encoder/include/constants.hpp:12)
constexpr uint32_t kSegmentUnits = 4;    // gpuhd SUBSEQ_SIZE (gpuhd/src/demo.cc:16)
constexpr uint32_t kTileThreads = 256;   // segments per tile (one per thread)
constexpr uint32_t kMaxLutBits = 12;     // primary multi-symbol LUT index width
constexpr uint32_t kMaxCodeLen = 16;     // .huff MAX_CODEWORD_LENGTH (decoder/include/constants.hpp:5)

// ---- geometry of the synchronisation kernels (K1a/K1b) --------------------
constexpr uint32_t kSyncChunkSegs = 16;  // segments per sync thread (2048 bits = two gap words)
constexpr uint32_t kSyncHaloSegs = 4;    // segments parsed speculatively BEFORE the chunk (512 bits; 6 measured: no gain)
constexpr uint32_t kSyncWarps = 23;      // warps per persistent sync CTA (one CTA per SM)
// A warp tile stages 32 chunks but OWNS only 31: lane 0 re-parses the last chunk of the previous tile (whose owner
// is lane 31 there) just to hand lane 1 an exit offset that had a whole chunk to synchronise -- so the seams between
// warp tiles are right without any communication between warps (3 % redundant parsing).
constexpr uint32_t kSyncTileChunks = 31;                                  // chunks owned per warp tile
constexpr uint32_t kSyncTileSegs = 32 * kSyncChunkSegs;                   // segments staged per warp tile
// Staged layout of a warp tile, in 16-byte units (= segments): [4 halo units][copy of chunk 0's first unit] then one
// REGION of 17 units per lane: [the lane's 16 segments][copy of the next chunk's first unit].  Every run a lane
// parses -- its halo (the tail of the previous region, or the tile halo) and its chunk -- is followed by the unit
// that comes next in the stream, so the flattened parse reads its words linearly; and the region stride of 17 units
// keeps 128-bit loads of the repair path conflict free.  Behind the regions: the first unit of the chunk after the
// tile + 2 units that the repair path may read but never uses, then the landing bytes (17 slots x 32 lanes).
constexpr uint32_t kSyncRegionUnits = kSyncChunkSegs + 1;
constexpr uint32_t kSyncTileSrcUnits = kSyncHaloSegs + kSyncTileSegs + 1;            // units copied per warp tile
constexpr uint32_t kSyncTilePhysUnits = kSyncHaloSegs + 1 + 32 * kSyncRegionUnits + 3;
constexpr uint32_t kSyncLandBytes = (kSyncChunkSegs + 1) * 32;
constexpr uint32_t kSyncTileBytes = kSyncTilePhysUnits * 16 + kSyncLandBytes;
// physical unit of tile-relative source unit i (i = 0: first halo unit)
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr uint32_t sync_phys_unit(uint32_t i) {
    return i < kSyncHaloSegs ? i
                             : kSyncHaloSegs + 1 + kSyncRegionUnits * ((i - kSyncHaloSegs) / kSyncChunkSegs) +
                                   ((i - kSyncHaloSegs) % kSyncChunkSegs);
}

// Primary LUT entry: byte 0 = total length in bits of the entry's codewords (1..16: bits 5-7 are zero, so the ENTRY
// ITSELF serves as a wrapping funnel-shift amount), byte 1 = first symbol, byte 2 = second symbol (0 if there is
// none), byte 3 = 8 * number of symbols.  The byte fields are added up with dp4a (FMA pipe) instead of being
// extracted with shifts (ALU pipe: the pipe that bounds the decode loop).  0 = "ask resolve_long" (LONG tables only).
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr uint32_t lut_entry(uint32_t len, uint32_t sym0, uint32_t sym1, uint32_t count) {
    return len | (sym0 << 8) | (sym1 << 16) | (count << 27);
}

// Host description of a prefix code: code/len indexed by symbol value.
struct HostCode {
    uint32_t code[256];
    uint8_t len[256];
    uint32_t n_present = 0;
    uint32_t min_len = 0;
    uint32_t max_len = 0;
    // symbols in canonical order (non-decreasing length, then increasing code)
    uint8_t order[256];
};

bool code_from_flat_lut(const hd_lut_entry *lut, uint32_t lut_bits, HostCode &c);
bool code_from_pairs(const hd_sym_len *pairs, uint32_t n, HostCode &c);

// Long-code (len > lut_bits) resolver tables, canonical first-code form.
struct LongTables {
    uint32_t first[17];   // first code of each length (right aligned)
    uint32_t count[17];   // number of codes of each length
    uint32_t offset[17];  // index of first symbol of each length in `order`
};

// All tables the kernels consume, laid out in ONE device blob:
//   [0]                      u32 dec_lut[1<<lut_bits]   lut_entry(): {bits, sym0, sym1, 8 * count}
//   [off_sync]               u16 sync_lut[1<<lut_bits]  {jump_bits, first_len}
//   [off_len]                u8  len_tab[256]
//   [off_order]              u8  order[256]
//   [off_long]               LongTables
struct TableBlob {
    uint32_t lut_bits = 0;
    uint32_t min_len = 0, max_len = 0;
    uint32_t has_long = 0;
    uint32_t off_sync = 0, off_len = 0, off_order = 0, off_long = 0;
    uint32_t total_bytes = 0;
    std::vector<uint8_t> bytes;
};

void build_table_blob(const HostCode &c, TableBlob &blob);

}  // namespace hd

// The opaque handle of the C ABI.
struct hd_table {
    hd::TableBlob host;
    uint8_t *d_blob = nullptr;
    int device = -1;
};

#endif  // HD_INTERNAL_H_
