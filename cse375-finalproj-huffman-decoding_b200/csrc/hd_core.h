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

#ifndef HD_TWO_STEP
#define HD_TWO_STEP 1   // decode_chunk: two table steps per window slide (measured)
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
// synthesized primary entry {sym, 0, bits=len, count=1}, or a 1-bit skip for a
// window that matches no code (hole of an incomplete code).
HD_HD uint32_t resolve_long(uint32_t win, uint32_t lut_bits, const LongTables *lt, const uint8_t *order) {
    for (uint32_t len = lut_bits + 1; len <= kMaxCodeLen; ++len) {
        const uint32_t c = win >> (32 - len);
        const uint32_t d = c - lt->first[len];
        if (d < lt->count[len]) return (uint32_t)order[lt->offset[len] + d] | (len << 16) | (1u << 24);
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
// transposed input tile: word w at in_col[w * NT] (conflict-free
// refills); words past the chunk (the next chunk's first two) must be present.
// Decoding starts at bit `entry` (< 32) and runs until the position reaches
// `stop` = chunk bits + entry offset of the NEXT chunk, i.e. exactly to the
// codeword boundary at which the successor starts (with an exact gap array this
// is the stop rule of decoder.cu:672 -- decode every codeword that STARTS inside
// the chunk -- because the last such codeword ends at that boundary).
//
// Symbols are stored one byte at a time into the thread's staging column:
// symbol i at row i/4, byte i%4 of the column.  The running store address A
// keeps the column bits; OR-ing them to ones before an add lets the carry out of
// the two byte bits ripple through the column field into the row field, so
// "advance by n symbols" is one add plus two logic ops.  A two-symbol LUT entry
// always stores both bytes; the second is simply overwritten by the next lookup
// when the entry held one symbol.  Returns the number of symbols.
template <int NT, bool LONG>
HD_HD uint32_t decode_chunk(const uint32_t *in_col, uint32_t entry, uint32_t stop, const DecodeTables &T,
                            uint32_t stage_addr, uint8_t *arena) {
    constexpr uint32_t kColBits = col_bits(NT);
    static_assert(row_shift(NT) != 0, "unsupported tile width");
    uint32_t w0 = in_col[0], w1 = in_col[NT];
    uint32_t w2 = in_col[2 * NT];
#if defined(__CUDA_ARCH__)
    uint32_t nxt = (uint32_t)__cvta_generic_to_shared(in_col + 2 * NT);   // shared-window address of the word in w2
#else
    const uint32_t *nxt = in_col + 2 * NT;
#endif
    const uint32_t col = stage_addr & kColBits;
    uint32_t A = stage_addr;
    uint32_t pos = entry;
    uint32_t e = 0;
#if defined(__CUDA_ARCH__)
#define HD_PLUS1(u_) add_fma((u_), T.one, 1u)
#else
#define HD_PLUS1(u_) ((u_) + 1u)
#endif
    while (pos < stop) {
        const uint32_t win = funnel_l(w1, w0, pos);   // next 32 stream bits
        e = T.lut[win >> T.idx_shift];
        if (LONG) {
            if ((e >> 16) == 0) e = resolve_long(win, T.lut_bits, T.lt, T.order);
        }
        uint32_t u = A | kColBits;
#if !(HD_WHATIF & 2)
        smem_st_u8(A, e, arena);
        smem_st_u8((HD_PLUS1(u) & ~kColBits) | col, e >> 8, arena);
#endif
        A = ((u + (e >> 24)) & ~kColBits) | col;
        uint32_t np = pos + byte2(e);
#if HD_TWO_STEP
        // Second table step from the same 32-bit window: a step consumes at most 16 bits, so the 16 bits the next one
        // looks at are still there (win << nbits).  One window slide, loop test and branch per TWO lookups, and the
        // loop-carried chain between the two loads is shift -> index -> load instead of add -> slide -> funnel -> index
        // -> load.  The step is dropped when the first one already reached `stop`.
        {
            const uint32_t win2 = win << byte2(e);
            uint32_t e2 = T.lut[win2 >> T.idx_shift];
            if (LONG) {
                if ((e2 >> 16) == 0) e2 = resolve_long(win2, T.lut_bits, T.lt, T.order);
            }
            if (np < stop) {
                e = e2;
#if defined(__CUDA_ARCH__)
                asm volatile("");
#endif
            } else {
                e2 = 0;
            }
            u = A | kColBits;
#if !(HD_WHATIF & 2)
            smem_st_u8(A, e2, arena);
            smem_st_u8((HD_PLUS1(u) & ~kColBits) | col, e2 >> 8, arena);
#endif
            A = ((u + (e2 >> 24)) & ~kColBits) | col;
            np += byte2(e2);
        }
#endif
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
#undef HD_PLUS1
    uint32_t cnt = (((A - stage_addr) >> row_shift(NT)) << 2) | (A & 3u);
    // the second symbol of the last lookup starts exactly at `stop`: it is the successor's first symbol
    if (pos > stop && (e >> 24) == 2u) --cnt;
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

// Copy c symbols from the thread's staging column (word m at src[m * stride],
// little-endian bytes) to byte offset D of the CTA's output tile: up to three
// head bytes until the destination is word aligned, whole 32-bit words through
// one funnel shift each, up to three tail bytes.  The words are fetched four at
// a time BEFORE the stores of the batch (source and destination are both shared
// memory, the compiler cannot reorder them itself), so that four loads are in
// flight instead of one load -> shift -> store chain per word; head and tail are
// predicated byte stores, not loops.  Reads at most three staging words past the
// last one that holds symbols (rows of the next buffer: harmless).
HD_HD void compact_to_tile(const uint32_t *src, uint32_t stride, uint32_t c, uint8_t *tile, uint32_t D) {
    uint32_t w = src[0];
    uint32_t head = (4u - (D & 3u)) & 3u;
    if (head > c) head = c;
    if (head > 0u) tile[D] = (uint8_t)w;
    if (head > 1u) tile[D + 1] = (uint8_t)(w >> 8);
    if (head > 2u) tile[D + 2] = (uint8_t)(w >> 16);
    const uint32_t rem = c - head;
    const uint32_t nw = rem >> 2;
    uint32_t *dst = reinterpret_cast<uint32_t *>(tile + D + head);
    const uint32_t sh = 8u * head;
    const uint32_t *nx = src + stride;
    for (uint32_t m = 0; m < nw; m += 4u) {
        const uint32_t n0 = nx[0], n1 = nx[stride], n2 = nx[2u * stride], n3 = nx[3u * stride];
        dst[0] = funnel_r(w, n0, sh);
        if (m + 1u < nw) dst[1] = funnel_r(n0, n1, sh);
        if (m + 2u < nw) dst[2] = funnel_r(n1, n2, sh);
        if (m + 3u < nw) dst[3] = funnel_r(n2, n3, sh);
        w = n3;
        nx += 4u * stride;
        dst += 4;
    }
    const uint32_t tail = rem & 3u;
    if (tail) {
        const uint32_t *lw = src + nw * stride;
        const uint32_t v = funnel_r(lw[0], lw[stride], sh);
        uint8_t *d8 = tile + D + head + 4u * nw;
        d8[0] = (uint8_t)v;
        if (tail > 1u) d8[1] = (uint8_t)(v >> 8);
        if (tail > 2u) d8[2] = (uint8_t)(v >> 16);
    }
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
