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
// synthesized primary entry (one symbol), or a 1-bit skip without symbols for a
// window that matches no code (hole of an incomplete code).
HD_HD uint32_t resolve_long(uint32_t win, uint32_t lut_bits, const LongTables *lt, const uint8_t *order) {
    for (uint32_t len = lut_bits + 1; len <= kMaxCodeLen; ++len) {
        const uint32_t c = win >> (32 - len);
        const uint32_t d = c - lt->first[len];
        if (d < lt->count[len]) return lut_entry(len, order[lt->offset[len] + d], 0u, 1u);
    }
    return lut_entry(1u, 0u, 0u, 0u);
}

struct DecodeTables {
    const uint32_t *lut;     // primary multi-symbol LUT {sym0, sym1, bits, count}
    uint32_t idx_shift;      // 32 - lut_bits
    uint32_t lut_bits;
    uint32_t one;            // the constant 1, opaque to the compiler (see decode_chunk)
    uint32_t sel_len, sel_cnt8, pair_slope;   // kSelLen, kSelCnt8, kPairSelSlope in registers (not re-materialised per iteration)
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

// c + byte k of e (k given as the dp4a selector 1 << 8k): ONE instruction on the FMA pipe (IDP.4A)
HD_HD uint32_t add_byte(uint32_t e, uint32_t sel, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp4a(e, sel, c);
#else
    uint32_t r = c;
    for (int k = 0; k < 4; ++k) r += ((e >> (8 * k)) & 0xffu) * ((sel >> (8 * k)) & 0xffu);
    return r;
#endif
}
HD_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    uint32_t d;   // (PTX directly: __byte_perm masks the selector first, one more ALU instruction)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7u))) & 0xffu) << (8 * k);
    return r;
#endif
}
constexpr uint32_t kSelLen = 0x00000001u, kSelCnt8 = 0x01000000u;   // dp4a selectors of the entry's length / 8 * count bytes

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
// the two words past the chunk (the next chunk's first two) must be present.
// Decoding starts at bit `entry` (< 32) and runs until the position reaches
// `stop` = chunk bits + entry offset of the NEXT chunk, i.e. exactly to the
// codeword boundary at which the successor starts (with an exact gap array this
// is the stop rule of decoder.cu:672 -- decode every codeword that STARTS inside
// the chunk -- because the last such codeword ends at that boundary).
//
// Two table steps per window slide: a step consumes at most 16 bits, so the 32-bit
// window taken at `pos` still holds the 16 bits the second step looks at.  BOTH
// steps always run -- the loop has no "did the first step reach stop?" test; what a
// pair decodes past `stop` (at most the second symbol of the step that reaches it
// and the whole step after it) is taken off the count after the loop and its bytes
// land in staging slack.  The <= 4 symbols of a pair are merged with one byte
// permute, shifted behind the bytes still pending in a register (P; n8 = 8 * symbols
// so far, used modulo 32 by wrapping shifts) and a full 32-bit word is stored to the
// thread's staging column (word m at stage_addr + m * NT * 4: bank == lane, conflict
// free) whenever four symbols are complete.  The loop is bound by the ALU pipe
// (shifts, logic, permutes: one warp instruction per 2 cycles); everything that can
// be an add goes to the FMA pipe instead (dp4a byte adds, multiply-adds, predicated
// moves).  Returns the number of symbols; staging words [0, ceil((count + 3) / 4)]
// may be written.
constexpr uint32_t kPairSel2 = 0x6521u;   // first step has two symbols: {e1.s0, e1.s1, e2.s0, e2.s1}
constexpr uint32_t kPairSel1 = 0x2651u;   // one symbol: {e1.s0, e2.s0, e2.s1, e1.s1 (= 0)}
constexpr uint32_t kPairSelSlope = (kPairSel2 - kPairSel1) / 8u;   // selector = base + slope * (8 * count): 8 -> Sel1, 16 -> Sel2
constexpr uint32_t kPairSelBase = kPairSel1 - 8u * kPairSelSlope;
static_assert(kPairSelBase + 16u * kPairSelSlope == kPairSel2, "pair selector is linear in the count byte");

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
    uint32_t P = 0, n8 = 0;    // pending symbols (low (n8 & 31) / 8 bytes of P); n8 = 8 * symbols so far
    uint32_t pos = entry;
    if (pos >= stop) {   // (an empty chunk: still leave a defined pending word)
        smem_st_u32(R, 0u, arena);
        return 0u;
    }
    uint32_t np1, e1, e2;
    do {
        const uint32_t win = funnel_l(w1, w0, pos);   // next 32 stream bits
        e1 = T.lut[win >> T.idx_shift];
        if (LONG) {
            if (e1 == 0) e1 = resolve_long(win, T.lut_bits, T.lt, T.order);
        }
        const uint32_t win2 = funnel_l(0u, win, e1);  // win << length of step 1 (the entry's low 5 bits)
        e2 = T.lut[win2 >> T.idx_shift];
        if (LONG) {
            if (e2 == 0) e2 = resolve_long(win2, T.lut_bits, T.lt, T.order);
        }
        np1 = add_byte(e1, T.sel_len, pos);
        const uint32_t np = add_byte(e2, T.sel_len, np1);
        // pack: X = the symbols of both steps
        uint32_t X;
        if (LONG) {   // (a skip entry has no symbols: general form)
            X = ((e1 >> 8) & 0xffffu) | (((e2 >> 8) & 0xffffu) << (e1 >> 24));
        } else {
            X = byte_perm(e1, e2, (e1 >> 24) * T.pair_slope + kPairSelBase);
        }
        const uint32_t n8n = add_byte(e2, T.sel_cnt8, add_byte(e1, T.sel_cnt8, n8));
        P += funnel_l(0u, X, n8);                 // X << (n8 & 31); the bytes are disjoint: an add (FMA pipe), not an or
        const uint32_t hi = funnel_l(X, 0u, n8);  // the bytes of X that did not fit (0 when n8 & 31 == 0)
#if defined(__CUDA_ARCH__)
        // four symbols complete: store the word, keep the overflow (one predicate; the adds are multiply-adds by the
        // opaque 1: FMA pipe)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 x;\n\t"
            "xor.b32 x, %3, %4;\n\t"
            "and.b32 x, x, 32;\n\t"
            "setp.ne.u32 p, x, 0;\n\t"
            "@p st.shared.u32 [%0], %1;\n\t"
            "@p mad.lo.u32 %0, %5, %6, %0;\n\t"   // R += row bytes
            "@p mad.lo.u32 %1, %2, %5, 0;\n\t"    // P = hi
            "}"
            : "+r"(R), "+r"(P)
            : "r"(hi), "r"(n8n), "r"(n8), "r"(T.one), "n"((HD_WHATIF & 64) ? 4 : NT * 4)   // (64: lane-contiguous stores, timing experiment)
            : "memory");
#else
        if ((n8n ^ n8) & 32u) {
            smem_st_u32(R, P, arena);
            R += NT * 4u;
            P = hi;
        }
#endif
        n8 = n8n;
#if defined(__CUDA_ARCH__)
        // Crossed a unit boundary: slide the window and fetch the word after it.  One asm block with its own predicate,
        // so that the slide stays predicated moves (FMA pipe; the ALU pipe bounds this loop) and the load writes w2 IN
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
    } while (pos < stop);
    smem_st_u32(R, P, arena);   // the pending symbols (a word past the last full one: the staging rows allow for it)
    // Take off what the last pair decoded past `stop`: the whole second step if the first one already reached it, and
    // the second symbol of the step that reached it if that symbol STARTS at `stop` (it is the successor's first one).
    uint32_t last = e2, lastpos = pos;
    if (np1 >= stop) {
        n8 -= (e2 >> 24);
        last = e1;
        lastpos = np1;
    }
    uint32_t cnt = n8 >> 3;
    if (lastpos > stop && (last >> 24) == 16u) --cnt;
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
    uint32_t one = 1;        // the constant 1, opaque to the compiler where it comes from a kernel parameter (see decode_chunk)
    bool use8 = false;       // optional byte copy of the jumps (K1a, non-EXACT: see hd_sync_parse_kernel); a constant
    uint32_t jump8_hi = 0;   // per kernel.  jump8_hi = its shared-window address >> lm (it is aligned to its size)
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
    uint32_t ent;
#if defined(__CUDA_ARCH__)
    if (!EXACT && L.use8) {   // (a compile-time constant of the kernel)
        // address = table base | index in ONE funnel shift: (base >> lm : win) >> (32 - lm)
        asm("ld.shared.u8 %0, [%1];" : "=r"(ent) : "r"(__funnelshift_r(win, L.jump8_hi, L.idx_shift)));
    } else
#endif
        ent = L.lut[win >> L.idx_shift];
    if (LONG) {
        if (ent == 0) {
            const uint32_t first = resolve_long(win, L.lm, L.lt, L.order) & 0xffu;
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
#if defined(__CUDA_ARCH__)
        uint32_t np = p1;   // + s2 while still inside the segment: one predicated multiply-add (see sync_run)
        asm("{\n\t"
            ".reg .pred p;\n\t"
            "setp.lt.u32 p, %1, 128;\n\t"
            "@p mad.lo.u32 %0, %2, %3, %0;\n\t"
            "}"
            : "+r"(np)
            : "r"(p1), "r"(s2), "r"(L.one));
#else
        const uint32_t np = p1 < kSegmentBits ? p1 + s2 : p1;
#endif
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
#if defined(__CUDA_ARCH__)
    uint32_t nxt = (uint32_t)__cvta_generic_to_shared(words + 2);   // shared-window address of the word in w2
#else
    const uint32_t *nxt = words + 2;
#endif
#if defined(__CUDA_ARCH__)
    uint32_t lp = land_addr + ((pos >> 7) << 5);   // landing slot of the segment that holds `pos`
#endif
    while (pos < end_bits) {
        const uint32_t win = funnel_l(w1, w0, pos);
        const uint32_t s1 = sync_step<EXACT, LONG>(win, pos & (kSegmentBits - 1u), L);
        const uint32_t p1 = pos + s1;
        const uint32_t s2 = sync_step<EXACT, LONG>(win << s1, p1 & (kSegmentBits - 1u), L);
#if defined(__CUDA_ARCH__)
        uint32_t np = p1;   // + s2 unless the first step crossed a segment boundary: one predicated multiply-add
        asm("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 x;\n\t"
            "xor.b32 x, %1, %2;\n\t"
            "and.b32 x, x, 128;\n\t"
            "setp.ne.u32 p, x, 0;\n\t"
            "@!p mad.lo.u32 %0, %3, %4, %0;\n\t"
            "}"
            : "+r"(np)
            : "r"(p1), "r"(pos), "r"(s2), "r"(L.one));
#else
        const uint32_t np = ((p1 ^ pos) & kSegmentBits) ? p1 : p1 + s2;
#endif
#if defined(__CUDA_ARCH__)
        if (RECORD) {
            // store into the slot of the segment the iteration started in; the slot pointer moves on when the position
            // crosses into the next segment (a pair of steps is at most 32 bits: never two segments): one logic
            // instruction + a predicated multiply-add instead of shift, mask and add per iteration
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                ".reg .b32 x;\n\t"
                "st.shared.u8 [%0], %1;\n\t"
                "xor.b32 x, %1, %2;\n\t"
                "and.b32 x, x, 128;\n\t"
                "setp.ne.u32 p, x, 0;\n\t"
                "@p mad.lo.u32 %0, %3, 32, %0;\n\t"
                "}"
                : "+r"(lp)
                : "r"(np), "r"(pos), "r"(L.one)
                : "memory");
        }
#else
        if (RECORD) smem_st_u8(land_addr + ((pos >> 7) << 5), np, arena);
#endif
#if defined(__CUDA_ARCH__)
        // Crossed a unit boundary: slide the window and fetch the word after it, as in decode_chunk: one block with
        // its own predicate, the pointer bumped and the word loaded IN PLACE (written as C the compiler spends two
        // more predicated moves per iteration on temporaries).
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 x;\n\t"
            "xor.b32 x, %4, %5;\n\t"
            "and.b32 x, x, 32;\n\t"
            "setp.ne.u32 p, x, 0;\n\t"
            "@p mad.lo.u32 %0, %1, %6, 0;\n\t"   // w0 = w1   (multiply by an opaque 1: IMAD on the FMA pipe, not SEL)
            "@p mad.lo.u32 %1, %2, %6, 0;\n\t"   // w1 = w2
            "@p mad.lo.u32 %3, %6, 4, %3;\n\t"   // nxt += 4
            "@p ld.shared.u32 %2, [%3];\n\t"
            "}"
            : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(nxt)
            : "r"(np), "r"(pos), "r"(L.one));
#else
        if ((np ^ pos) & 32u) {
            w0 = w1;
            w1 = w2;
            ++nxt;
            w2 = *nxt;
        }
#endif
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
