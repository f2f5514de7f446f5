// hd_core.h -- per-thread building blocks of the decode kernels, written as
// host/device inline functions so that the SAME source can be single-stepped by
// the CPU emulator under tests/emul (test infrastructure; the product library
// only ever runs them inside the CUDA kernels of hd_kernels.cuh).
#ifndef HD_CORE_H_
#define HD_CORE_H_

#include <stdint.h>
#include <string.h>

#include "hd_internal.h"

#if defined(__CUDACC__)
#define HD_HD __host__ __device__ __forceinline__
#else
#define HD_HD inline
#endif

#ifndef HD_WHATIF
#define HD_WHATIF 0   // developer timing experiments (tools/whatif.sh); 0 in every shipped build
#endif

namespace hd {

// upper 32 bits of (hi:lo) << (s & 31)
HD_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? ((hi << s) | (lo >> (32u - s))) : hi;
#endif
}

// Canonical resolver for codes longer than the primary LUT (len in
// (lut_bits, 16]).  win = next 32 stream bits, left aligned.  Returns a
// synthesized primary entry {sym, 0, bits=len, 8 * count = 8}, or a 1-bit skip for a
// window that matches no code (hole of an incomplete code).
HD_HD uint32_t resolve_long(uint32_t win, uint32_t lut_bits, const LongTables *lt, const uint8_t *order) {
    for (uint32_t len = lut_bits + 1; len <= kMaxCodeLen; ++len) {
        const uint32_t c = win >> (32 - len);
        const uint32_t d = c - lt->first[len];
        if (d < lt->count[len]) return (uint32_t)order[lt->offset[len] + d] | (len << 16) | (8u << 24);
    }
    return 1u << 16;
}

struct DecodeTables {
    const uint32_t *lut;     // primary multi-symbol LUT {sym0, sym1, bits, count}
    uint32_t idx_shift;      // 32 - lut_bits
    uint32_t lut_bits;
    uint32_t one;            // the constant 1, opaque to the compiler (see decode_chunk)
    const uint8_t *len_tab;  // code length by symbol
    const LongTables *lt;
    const uint8_t *order;
};

// ---- "shared memory" byte/word accessors --------------------------------------
// On the device `a` is an absolute 32-bit shared-window address (so that address
// arithmetic needs no base add and maps 1:1 onto STS/LDS operands); in the CPU
// emulator it is an offset into a host arena that mimics the CTA's smem.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void smem_st_u8(uint32_t a, uint32_t v, uint8_t *) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ void smem_st_u32(uint32_t a, uint32_t v, uint8_t *) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ uint32_t smem_ld_u32(uint32_t a, const uint8_t *) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#else
inline void smem_st_u8(uint32_t a, uint32_t v, uint8_t *arena) { arena[a] = (uint8_t)v; }
inline void smem_st_u32(uint32_t a, uint32_t v, uint8_t *arena) { memcpy(arena + a, &v, 4); }
inline uint32_t smem_ld_u32(uint32_t a, const uint8_t *arena) { uint32_t v; memcpy(&v, arena + a, 4); return v; }
#endif

// Staging/input tiles are TRANSPOSED: with NT decode threads per tile, thread t
// owns byte lanes [4t, 4t+3] of every row of NT*4 bytes (bank == lane: conflict
// free).  col_bits(NT) = the address bits that select the thread's column.
#if defined(__CUDACC__)
#define HD_CONSTEXPR __host__ __device__ constexpr
#else
#define HD_CONSTEXPR constexpr
#endif
HD_CONSTEXPR uint32_t col_bits(uint32_t nt) { return (nt - 1u) << 2; }
HD_CONSTEXPR uint32_t row_shift(uint32_t nt) { return nt == 512 ? 11u : nt == 256 ? 10u : nt == 128 ? 9u : nt == 32 ? 7u : 0u; }

HD_HD uint32_t byte2(uint32_t e) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(e, 0u, 0x4442u);   // one PRMT instead of shift + mask
#else
    return (e >> 16) & 0xffu;
#endif
}

// a + k as a multiply-add with an OPAQUE multiplier register holding 1: it issues on the FMA pipe.  The decode loop is
// bound by the INT (ALU) pipe, which retires one warp instruction every two cycles; written as an add the compiler
// puts it there (VIADD).  (The same trick for the two shifts of a table step, mad.hi by an opaque power of two, was
// measured and is slower: IMAD.HI is not a cheap instruction.)
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t one, uint32_t k) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(one), "r"(k), "r"(a));
    return d;
}
#endif

// Decode one chunk (SEGS consecutive 128-bit segments of one thread) in a single
// flattened loop.  The chunk's units sit in the thread's column of the
// transposed input block: word w at in_col[w * NT] (conflict-free refills);
// words past the chunk (the next chunk's first two) must be present.
// Decoding starts at bit `entry` (< 32) and runs until the position reaches
// `stop` = chunk bits + entry offset of the NEXT chunk, i.e. exactly to the
// codeword boundary at which the successor starts (with an exact gap array this
// is the stop rule of decoder.cu:672 -- decode every codeword that STARTS inside
// the chunk -- because the last such codeword ends at that boundary).
//
// Two table steps per window slide: a step consumes at most 16 bits, so the 32-bit
// window taken at `pos` still holds the 16 bits the second step looks at
// (win << bits1); the second step is dropped when the first reached `stop`.  The
// <= 4 symbols of the two steps are packed behind the bytes still pending in a
// register (P, np8 = 8 * pending bytes) and a full 32-bit word is stored to the
// thread's staging column (word m at stage_addr + m * NT * 4: bank == lane, conflict
// free) whenever four symbols are complete: ONE shared-memory store per four
// symbols instead of one per symbol -- the kernel is bound by shared-memory
// wavefronts.  LUT entries are {sym0, sym1, bits, 8 * count}.
// Returns the number of symbols; staging words [0, ceil(count / 4)] are written.
template <int NT, bool LONG>
HD_HD uint32_t decode_chunk(const uint32_t *in_col, uint32_t entry, uint32_t stop, const DecodeTables &T,
                            uint32_t stage_addr, uint8_t *arena) {
    static_assert(row_shift(NT) != 0, "unsupported tile width");
    uint32_t w0 = in_col[0], w1 = in_col[NT];
    uint32_t w2 = in_col[2 * NT];
#if defined(__CUDA_ARCH__)
    uint32_t nxt = (uint32_t)__cvta_generic_to_shared(in_col + 2 * NT);   // shared-window address of the word in w2
#else
    const uint32_t *nxt = in_col + 2 * NT;
#endif
    uint32_t R = stage_addr;   // where the next full word goes
    uint32_t P = 0, np8 = 0;   // pending symbols (low np8 / 8 bytes of P), np8 < 32
    uint32_t pos = entry;
    uint32_t e = 0;
    while (pos < stop) {
        const uint32_t win = funnel_l(w1, w0, pos);   // next 32 stream bits
        uint32_t e1 = T.lut[win >> T.idx_shift];
        if (LONG) {
            if ((e1 >> 16) == 0) e1 = resolve_long(win, T.lut_bits, T.lt, T.order);
        }
        const uint32_t b1 = byte2(e1);
        uint32_t np = pos + b1;
        const uint32_t win2 = win << b1;
        uint32_t e2 = T.lut[win2 >> T.idx_shift];
        if (LONG) {
            if ((e2 >> 16) == 0) e2 = resolve_long(win2, T.lut_bits, T.lt, T.order);
        }
        if (np < stop) {
            e = e2;
        } else {
            e = e1;
            e2 = 0;
        }
        np += byte2(e2);
        // pack: X = the symbols of both steps, n8 = 8 * (pending + new) symbols
        const uint32_t c1 = e1 >> 24;
        const uint32_t X = (e1 & 0xffffu) | ((e2 & 0xffffu) << c1);
        const uint32_t n8 = np8 + c1 + (e2 >> 24);
        const uint32_t lo = P | (X << np8);
        const uint32_t hi = funnel_l(X, 0u, np8);   // the bytes of X that did not fit into lo (0 when np8 == 0)
        if (n8 >= 32u) {
            smem_st_u32(R, lo, arena);
            R += NT * 4u;
            P = hi;
        } else {
            P = lo;
        }
        np8 = n8 & 31u;
#if defined(__CUDA_ARCH__)
        // Crossed a unit boundary: slide the window and fetch the word after it.  One asm block with its own predicate,
        // so that the slide stays predicated moves (FMA pipe; the INT pipe bounds this loop) and the load writes w2 IN
        // PLACE: it is consumed by the next slide, at least two lookups later, so its latency never sits on the
        // loop-carried path and it costs a shared-memory wavefront once per unit instead of once per lookup.
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 x;\n\t"
            "xor.b32 x, %4, %5;\n\t"
            "and.b32 x, x, 32;\n\t"
            "setp.ne.u32 p, x, 0;\n\t"
            "@p mad.lo.u32 %0, %1, %6, 0;\n\t"    // w0 = w1   (multiply by an opaque 1: IMAD, not SEL)
            "@p mad.lo.u32 %1, %2, %6, 0;\n\t"    // w1 = w2
            "@p mad.lo.u32 %3, %6, %7, %3;\n\t"   // nxt += row bytes
            "@p ld.shared.u32 %2, [%3];\n\t"
            "}"
            : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(nxt)
            : "r"(np), "r"(pos), "r"(T.one), "n"(NT * 4));
#else
        if ((np ^ pos) & 32u) {
            w0 = w1;
            w1 = w2;
            nxt += NT;
            w2 = *nxt;
        }
#endif
        pos = np;
    }
    smem_st_u32(R, P, arena);   // the pending symbols (a word past the last full one: the staging rows allow for it)
    uint32_t cnt = (((R - stage_addr) >> row_shift(NT)) << 2) + (np8 >> 3);
    // the second symbol of the last lookup starts exactly at `stop`: it is the successor's first symbol
    if (pos > stop && (e >> 24) == 16u) --cnt;
    return cnt;
}

// lower 32 bits of (hi:lo) >> (s & 31)
HD_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? ((lo >> s) | (hi << (32u - s))) : lo;
#endif
}

// ---- warp-private staging blocks (K4 "warp tile" kernel) ------------------------
// A decode warp owns a block of `rows` rows of 32 words: lane l's symbols sit in
// column l (word m of the lane at word index m * 32 + l, little-endian bytes) --
// the layout decode_chunk<32> writes.  The block is compacted IN PLACE: every
// lane first pulls its whole column into registers (column_read), the warp
// synchronises, then every lane writes its c symbols to byte offset D of the
// same block (column_write), D = 16-byte phase of the warp's global destination
// + the exclusive prefix of the lanes' counts -- after which the warp's span is
// one contiguous run that a TMA bulk store flushes.  No separate output tile.
//
// column_write stores whole 32-bit words (one funnel shift each) from the first
// word-aligned byte on, INCLUDING a last partial word: its surplus bytes land on
// the head bytes of the following lane(s), which is why the <= 3 head bytes of
// every lane are written by column_write_head AFTER a second warp
// synchronisation.  (A lane whose symbols all fit before the first word boundary
// writes nothing but head bytes.)  Surplus bytes of the last lane fall into the
// slack behind the warp's span.  w[] must hold RCAP + 2 words; words past the
// lane's rows may be garbage (they only reach bytes nobody keeps).
// Addresses are shared-window addresses on the device (explicit LDS/STS, no generic-pointer conversion per access)
// and arena offsets in the CPU emulator, as for decode_chunk.
template <int RCAP>
HD_HD void column_read(uint32_t col_addr, uint32_t rows_max, uint32_t (&w)[RCAP + 2], const uint8_t *arena) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < RCAP; g += 4) {
        if ((uint32_t)g < rows_max) {   // warp-uniform: whole groups are skipped, not predicated off
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = g; i < g + 4 && i < RCAP; ++i) w[i] = smem_ld_u32(col_addr + (uint32_t)i * 128u, arena);
        }
    }
}

// D = byte address of the lane's first symbol in the compacted block
template <int RCAP>
HD_HD void column_write(const uint32_t (&w)[RCAP + 2], uint32_t c, uint32_t D, uint32_t rows_max, uint8_t *arena) {
    uint32_t head = (4u - (D & 3u)) & 3u;
    if (head > c) head = c;
    const uint32_t nw = (c - head + 3u) >> 2;   // words incl. the partial last one
    const uint32_t dst = D + head;
    const uint32_t sh = 8u * head;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < RCAP; g += 4) {
        if ((uint32_t)g < rows_max) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = g; i < g + 4 && i < RCAP; ++i)
                if ((uint32_t)i < nw) smem_st_u32(dst + 4u * (uint32_t)i, funnel_r(w[i], w[i + 1], sh), arena);
        }
    }
}

HD_HD void column_write_head(uint32_t w0, uint32_t c, uint32_t D, uint8_t *arena) {
    uint32_t head = (4u - (D & 3u)) & 3u;
    if (head > c) head = c;
    if (head > 0u) smem_st_u8(D, w0, arena);
    if (head > 1u) smem_st_u8(D + 1u, w0 >> 8, arena);
    if (head > 2u) smem_st_u8(D + 2u, w0 >> 16, arena);
}

// ---------------------------------------------------------------- sync side
struct SyncLut {
    const uint16_t *lut;   // {jump_bits | first_len << 8}
    uint32_t idx_shift;
    uint32_t lm;
    const LongTables *lt;
    const uint8_t *order;
};

struct Seg4 {
    uint32_t x, y, z, w;
};

// Parse ONE 128-bit segment held in r0..r3 (+ r4 = first unit of the next
// segment) from bit `pos` (< 32) and return the offset, relative to the start
// of the NEXT segment, at which the parse lands there ("landing", 0..15).
//
// The sync LUT entry holds {jump, first}: jump = total length of all complete
// codewords inside the LUT window, first = length of the first one.
//   EXACT = false: always take the whole jump.  The landing is then a codeword
//     boundary of this parse at or after the segment end, but not necessarily
//     the FIRST one (a jump may skip over a codeword that starts exactly at the
//     boundary).  Because every parse uses the same table, two parses that have
//     merged still produce identical landings, which is all synchronisation
//     needs; and the decode kernel stops a chunk exactly at the landing recorded
//     for its successor, so such a gap array decodes to the same bytes.
//   EXACT = true: a jump that would pass the segment end is replaced by single
//     codeword steps, so that the landing is the offset of the FIRST codeword
//     starting in the next segment -- the .huff gap-array definition
//     (encoder.cu:307-312), needed by hd_selfsync_gap_array.
// Merge test of the reference: cuhd_gpu_decoder.cu:208-213.
template <bool EXACT, bool LONG>
HD_HD uint32_t sync_step(uint32_t win, uint32_t pos, const SyncLut &L) {
    uint32_t ent = L.lut[win >> L.idx_shift];
    if (LONG) {
        if (ent == 0) {
            const uint32_t first = (resolve_long(win, L.lm, L.lt, L.order) >> 16) & 0xffu;
            ent = first | (first << 8);
        }
    }
    uint32_t step = ent & 0xffu;
    if (EXACT) {
        if (pos + step > kSegmentBits) step = ent >> 8;
    }
    return step;
}

template <bool EXACT, bool LONG>
HD_HD uint32_t sync_segment(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t r4, uint32_t pos,
                            const SyncLut &L) {
    // Two table steps per window slide: a step is at most 16 bits, so the 32-bit window taken at `pos` still holds
    // the 16 bits the second step looks at (hi << step1), and the pair crosses at most one unit boundary.  The second
    // step is dropped when the first already left the segment, so the sequence of positions -- and therefore every
    // landing -- is exactly that of single steps.
    while (pos < kSegmentBits) {
        const uint32_t win = funnel_l(r1, r0, pos);
        const uint32_t s1 = sync_step<EXACT, LONG>(win, pos, L);
        const uint32_t p1 = pos + s1;
        const uint32_t s2 = sync_step<EXACT, LONG>(win << s1, p1, L);
        const uint32_t np = p1 < kSegmentBits ? p1 + s2 : p1;
        if ((np ^ pos) & 32u) {
            r0 = r1; r1 = r2; r2 = r3; r3 = r4;
#if defined(__CUDA_ARCH__)
            asm volatile("");   // keep predicated moves (FMA pipe) instead of selects (INT pipe), see decode_chunk
#endif
        }
        pos = np;
    }
    return pos - kSegmentBits;
}

// Flattened parse of a RUN of segments (K1a): from bit `pos` of words[0] until the position reaches end_bits, in
// one loop -- no reconvergence at segment ends (with one loop per segment every lane of a warp waits for the
// slowest lane at each of its 20 segments: ~40 % of the issue slots of the round-1 kernel).  The sequence of
// positions is exactly that of sync_segment called segment by segment: two table steps per iteration, the second
// one dropped when the first crosses a 128-bit boundary, so every landing is identical.
//
// RECORD: the landing into segment s+1 is left in byte slot s+1 of the lane's landing column (slot k at land_addr
// + 32 * k: one byte per lane and row).  Every iteration stores the low byte of its end position into the slot
// FOLLOWING the segment it started in; the last iteration that starts in segment s is the one that crosses into
// s+1, so the slot ends up holding (landing + a multiple of 128) & 0xff -- an unconditional store (two
// instructions) instead of a divergent "did I cross?" block.  Words are read linearly from `words` (the run and
// at least 64 bits behind it must be readable).  Returns the final position (>= end_bits).
template <bool EXACT, bool LONG, bool RECORD>
HD_HD uint32_t sync_run(const uint32_t *words, uint32_t pos, uint32_t end_bits, const SyncLut &L, uint32_t land_addr,
                        uint8_t *arena) {
    uint32_t w0 = words[0], w1 = words[1], w2 = words[2];
    const uint32_t *nxt = words + 2;
    while (pos < end_bits) {
        const uint32_t win = funnel_l(w1, w0, pos);
        const uint32_t s1 = sync_step<EXACT, LONG>(win, pos & (kSegmentBits - 1u), L);
        const uint32_t p1 = pos + s1;
        const uint32_t s2 = sync_step<EXACT, LONG>(win << s1, p1 & (kSegmentBits - 1u), L);
        const uint32_t np = ((p1 ^ pos) & kSegmentBits) ? p1 : p1 + s2;
        if (RECORD) smem_st_u8(land_addr + ((pos >> 7) << 5), np, arena);
        if ((np ^ pos) & 32u) {
            w0 = w1;
            w1 = w2;
            ++nxt;
            w2 = *nxt;
#if defined(__CUDA_ARCH__)
            asm volatile("");   // keep predicated moves (FMA pipe) instead of selects (INT pipe), see decode_chunk
#endif
        }
        pos = np;
    }
    return pos;
}

// Repair step (K1b): re-parse chunk `chunk` (kSyncChunkSegs segments) from the
// exact entry offset `entry` and overwrite the recorded landings until the new
// parse lands where the recorded one did (the parses have merged and are
// identical from there on).  words[0..1] = the chunk's two gap words.  Returns
// true if the parses merged inside the chunk; false means the chunk's exit
// changed and the next chunk has to be looked at again.
template <bool EXACT, bool LONG, typename LoadSeg>
HD_HD bool sync_repair_chunk(LoadSeg load_seg, uint64_t chunk, uint32_t entry, const SyncLut &L, uint32_t words[2]) {
    const uint64_t seg0 = chunk * kSyncChunkSegs;
    Seg4 cur = load_seg(seg0);
    Seg4 nxt = load_seg(seg0 + 1);
    uint32_t pos = entry;
    for (uint32_t k = 0; k < kSyncChunkSegs; ++k) {
        pos = sync_segment<EXACT, LONG>(cur.x, cur.y, cur.z, cur.w, nxt.x, pos, L);
        const uint32_t shift = 4u * (k & 7u);
        uint32_t &word = words[k >> 3];
        if (((word >> shift) & 15u) == pos) return true;
        word = (word & ~(15u << shift)) | (pos << shift);
        cur = nxt;
        nxt = load_seg(seg0 + k + 2);
    }
    return false;
}

}  // namespace hd
#endif  // HD_CORE_H_
