/*
 * hd_b200.h -- C ABI of the B200-native (sm_100a) parallel Huffman decoder.
 *
 * Drop-in boundary for ONE hot path of dek226/CSE375-FinalProj-Huffman-Decoding:
 * decoding a canonical Huffman bitstream in parallel through
 *   (a) the self-synchronisation scheme  (gpuhd/src/cuhd_gpu_decoder.cu) and
 *   (b) the gap-array scheme  (Huffman_coding_Gap_arrays/decoder/src/decoder.cu).
 * Plain pointers and sizes only; no C++/torch types.  Every entry point cites
 * the reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *  - "units" are host-endian uint32 words, bit 31 of unit 0 is the first code
 *    bit (gpuhd/encoder/src/llhuffman_encoder.cc:216-231,
 *    Huffman_coding_Gap_arrays/encoder/src/encoder.cu:295-331).
 *  - d_* pointers are device pointers on the current CUDA device; `stream` is a
 *    cudaStream_t passed as void* (NULL = legacy default stream).
 *  - All functions return hd_status instead of the reference's print-and-exit
 *    (gpuhd/include/cuhd_cuda_definitions.h:21-28, decoder.cu:14-20); the C++
 *    shim classes keep exit(1) for the demos.
 *  - No CPU fallback exists: without a CUDA device every compute call returns
 *    HD_ERR_CUDA.
 */
#ifndef HD_B200_H_
#define HD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    HD_OK = 0,
    HD_ERR_INVALID = 1,     /* bad argument */
    HD_ERR_CUDA = 2,        /* CUDA runtime error (see hd_last_cuda_error) */
    HD_ERR_WORKSPACE = 3,   /* workspace too small */
    HD_ERR_FORMAT = 4,      /* malformed .huff image / code table */
    HD_ERR_UNRESOLVED = 5   /* self-sync fixed point not reached (internal) */
} hd_status;

/* == cuhd::CUHDCodetableItemSingle  (gpuhd/include/cuhd_codetable.h:20-23) */
typedef struct {
    uint8_t num_bits;
    uint8_t symbol;
} hd_lut_entry;

/* one {symbol,length} pair of a .huff header
 * (Huffman_coding_Gap_arrays/encoder/src/huff.cpp:189-194) */
typedef struct {
    uint8_t symbol;
    uint8_t length;
} hd_sym_len;

const char *hd_version(void);
const char *hd_status_string(hd_status s);
/* last cudaError_t seen by this library on the calling thread, as text */
const char *hd_last_cuda_error(void);
/* optional: create the CUDA context of the current device, load the decode
 * kernels and make the calling thread's pinned convergence flag now (CUDA loads
 * kernel code lazily at the first launch and a pinned allocation costs
 * milliseconds; the reference demos time their first and only decode,
 * gpuhd/src/demo.cc:170-197).  Call it on the thread that will decode. */
hd_status hd_warmup(void);

/* ------------------------------------------------------------------ tables
 * Device-resident decode tables.  Replaces cuhd::CUHDGPUCodetable
 * (gpuhd/include/cuhd_gpu_codetable.h, allocate + cpy_host_to_device) and the
 * two-level table of get_table_info/get_twolevel_table
 * (Huffman_coding_Gap_arrays/decoder/src/get_table.cpp:3-139). */
typedef struct hd_table hd_table;

/* From the reference's flat LUT: 2^lut_bits x {num_bits,symbol}, index = next
 * lut_bits stream bits (gpuhd/encoder/src/llhuffman_encoder.cc:240-262).
 * lut_bits <= 16. */
hd_status hd_table_from_flat_lut(const hd_lut_entry *lut, uint32_t lut_bits,
                                 void *stream, hd_table **out);
/* From a .huff header list in file order (non-decreasing length, canonical
 * codes by code=(code+1)<<(next_len-len); decoder/src/get_table.cpp:17-20).
 * Lengths 1..16. */
hd_status hd_table_from_lengths(const hd_sym_len *pairs, uint32_t n_pairs,
                                void *stream, hd_table **out);
/* Same tables built ON THE DEVICE by a kernel from a device-resident list
 * (e.g. the header of a .huff image that is already in HBM): no host-side
 * table construction and no table upload; only a 20-byte summary is read back.
 * Replaces get_table_info/get_twolevel_table (decoder/src/get_table.cpp:3-139)
 * and get_decoder_table (llhuffman_encoder.cc:240-262).  Synchronises `stream`. */
hd_status hd_table_from_lengths_device(const hd_sym_len *d_pairs, uint32_t n_pairs,
                                       void *stream, hd_table **out);
/* copies the table's device blob to host memory (dst may be NULL to query the
 * size); for tests and debugging */
hd_status hd_table_blob(const hd_table *t, void *dst, size_t cap, size_t *size);
void hd_table_free(hd_table *t);
uint32_t hd_table_max_length(const hd_table *t);
uint32_t hd_table_min_length(const hd_table *t);

/* --------------------------------------------------------------- workspace
 * Replaces cuhd::CUHDGPUDecoderMemory (gpuhd/src/cuhd_gpu_decoder_memory.cc:
 * 25-48: 16 B sync point + 4 B count per subsequence + flags) and the
 * per-segment output_pos array of decoder.cu:754.  Ours: one 64-bit status word,
 * padded to a 32-byte sector, per 256 segments (an upper bound: the kernel's
 * tiles hold 512-1728 segments) + (self-sync) 4 bits per 128-bit subsequence,
 * one byte per 16 subsequences and one word per 496 -- about 0.035 B per
 * stream byte. */
size_t hd_workspace_bytes_gap(uint64_t n_units);
size_t hd_workspace_bytes_selfsync(uint64_t n_units);

/* ------------------------------------------------------------------ decode
 * Gap-array scheme.  Replaces gpu_dec_l1_l2 + decoder_l1_l2
 * (Huffman_coding_Gap_arrays/decoder/src/decoder.cu:454-730, 732-815;
 * declared decoder/include/decoder.cuh:4-15).  d_gap_words: ceil(n_segments/8)
 * words, element i in word i/8 bits [4(i%8),4(i%8)+3]; segment k>0 starts at
 * bit 128k+gap[k-1] (decoder.cu:505-507).  Element n_segments-1 is read too
 * (the offset at which the parse leaves the stream); d_gap_words may be NULL
 * only when n_segments == 1.  Writes exactly n_symbols bytes (fewer only if the
 * stream holds fewer codewords).  Asynchronous on stream. */
hd_status hd_decode_gap(const uint32_t *d_units, uint64_t n_units,
                        const uint32_t *d_gap_words, uint64_t n_segments,
                        const hd_table *table, uint8_t *d_out,
                        uint64_t n_symbols, void *d_workspace,
                        size_t workspace_bytes, void *stream);

/* Self-synchronisation scheme.  Replaces cuhd::CUHDGPUDecoder::decode
 * (gpuhd/src/cuhd_gpu_decoder.cu:422-523; declared
 * gpuhd/include/cuhd_gpu_decoder.h:24-32) = phases 1-4.  n_units may include
 * the reference's trailing zero pad unit (gpuhd/src/cuhd_input_buffer.cc:
 * 17-31).  The stream must start at bit 0 of unit 0 unless first_bit is set
 * through hd_decode_selfsync_ex.  Synchronises the stream once (reads one
 * 4-byte convergence flag), like the reference's phase-2 loop (:461-495). */
hd_status hd_decode_selfsync(const uint32_t *d_units, uint64_t n_units,
                             const hd_table *table, uint8_t *d_out,
                             uint64_t n_symbols, void *d_workspace,
                             size_t workspace_bytes, void *stream);

/* The same decode split in two for callers that overlap it with other work:
 * _enqueue puts everything (one memset, K1a, K4, the 4-byte copy of the
 * convergence counter into *h_unconverged -- pinned host memory) on
 * `stream` and returns without synchronising; _finish synchronises the stream
 * and, in the rare case that *h_unconverged != 0 (a parse that did not
 * synchronise within a 64 Kbit warp tile), iterates K1b to the fixed point
 * and decodes again.  hd_decode_selfsync[_ex] == _enqueue + _finish. */
hd_status hd_decode_selfsync_enqueue(const uint32_t *d_units, uint64_t n_units,
                                     uint32_t first_bit, const hd_table *table,
                                     uint8_t *d_out, uint64_t n_symbols,
                                     void *d_workspace, size_t workspace_bytes,
                                     uint64_t *d_count,
                                     uint32_t *h_unconverged, void *stream);
hd_status hd_decode_selfsync_finish(const uint32_t *d_units, uint64_t n_units,
                                    uint32_t first_bit, const hd_table *table,
                                    uint8_t *d_out, uint64_t n_symbols,
                                    void *d_workspace, size_t workspace_bytes,
                                    uint64_t *d_count,
                                    const uint32_t *h_unconverged, void *stream);

/* Extended entries used by the multi-GPU sharding layer (they replace the
 * slicing of gpuhd/multigpu_demo.cc:178-227): the shard's first segment starts
 * at bit `first_bit` (0..15) instead of 0, and *d_count (device uint64, may be
 * NULL) receives the number of symbols written -- the one value per GPU that
 * the NCCL all-gather exchanges to place the shards' outputs.  For the gap
 * entry n_units counts the units that may be READ and n_segments (<=
 * ceil(n_units/4)) the segments to decode: a shard that is not the last one
 * passes one extra unit, because its last codeword may end there. */
hd_status hd_decode_gap_ex(const uint32_t *d_units, uint64_t n_units,
                           const uint32_t *d_gap_words, uint64_t n_segments,
                           uint32_t first_bit, const hd_table *table,
                           uint8_t *d_out, uint64_t n_symbols,
                           void *d_workspace, size_t workspace_bytes,
                           uint64_t *d_count, void *stream);
hd_status hd_decode_selfsync_ex(const uint32_t *d_units, uint64_t n_units,
                                uint32_t first_bit, const hd_table *table,
                                uint8_t *d_out, uint64_t n_symbols,
                                void *d_workspace, size_t workspace_bytes,
                                uint64_t *d_count, void *stream);

/* One shard of a longer stream, for multi-GPU decode (replaces the per-GPU
 * slicing of gpuhd/multigpu_demo.cc:186-199, which cuts units and symbols
 * independently and therefore fails).  d_units -> [left halo | shard | right
 * halo]: n_units_halo (multiple of 32) units before the shard are parsed
 * speculatively so that the parse is synchronised when it enters the shard;
 * units up to n_units_readable (>= halo+shard; the last codeword of the shard
 * may end in the next unit) may be read.  Rank 0 passes halo 0 and first_bit.
 * d_shard_info receives {symbols written, entry bit offset used, bit offset at
 * which the parse leaves the shard}: ranks all-gather these (NCCL), check
 * entry[g] == exit[g-1] and exclusive-scan the counts for output placement. */
hd_status hd_decode_selfsync_shard(const uint32_t *d_units, uint64_t n_units_halo,
                                   uint64_t n_units_shard,
                                   uint64_t n_units_readable, uint32_t first_bit,
                                   const hd_table *table, uint8_t *d_out,
                                   uint64_t out_cap, void *d_workspace,
                                   size_t workspace_bytes,
                                   uint64_t *d_shard_info, void *stream);

/* the same split in two, like hd_decode_selfsync_enqueue / _finish: _enqueue
 * only queues work on `stream` (so that the caller can queue its NCCL
 * all-gather of d_shard_info right behind it), _finish synchronises and
 * handles the rare unconverged case. */
hd_status hd_decode_selfsync_shard_enqueue(const uint32_t *d_units, uint64_t n_units_halo,
                                           uint64_t n_units_shard,
                                           uint64_t n_units_readable,
                                           uint32_t first_bit, const hd_table *table,
                                           uint8_t *d_out, uint64_t out_cap,
                                           void *d_workspace, size_t workspace_bytes,
                                           uint64_t *d_shard_info,
                                           uint32_t *h_unconverged, void *stream);
hd_status hd_decode_selfsync_shard_finish(const uint32_t *d_units, uint64_t n_units_halo,
                                          uint64_t n_units_shard,
                                          uint64_t n_units_readable,
                                          uint32_t first_bit, const hd_table *table,
                                          uint8_t *d_out, uint64_t out_cap,
                                          void *d_workspace, size_t workspace_bytes,
                                          uint64_t *d_shard_info,
                                          const uint32_t *h_unconverged, void *stream);

/* Synchronisation only (phases 1-2 of the reference, cuhd_gpu_decoder.cu:
 * 145-327): writes the gap array of the stream -- for every 128-bit
 * subsequence k>0 the bit offset of the first codeword starting in it, packed
 * like the .huff gap array -- into d_gap_words_out (ceil(n_seg/8) words,
 * n_seg = ceil(n_units/4)).  After this, hd_decode_gap decodes the stream. */
hd_status hd_selfsync_gap_array(const uint32_t *d_units, uint64_t n_units,
                                uint32_t first_bit, const hd_table *table,
                                uint32_t *d_gap_words_out, void *d_workspace,
                                size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------ .huff format
 * Replaces the parser of Huffman_coding_Gap_arrays/decoder/src/huff.cpp:36-100
 * and the writer of encoder/src/huff.cpp:186-202.  Pointers in the view alias
 * the caller's buffer (payload may be 2-byte aligned only: copy before use as
 * uint32). */
typedef struct {
    uint64_t symbol_count;
    const hd_sym_len *symbols;   /* symbol_count pairs, file order */
    uint32_t original_size;      /* decoded bytes */
    uint32_t n_units;            /* compressed 32-bit units */
    uint32_t n_segments;         /* 128-bit segments = gap elements */
    const uint8_t *gap_words;    /* 4*ceil(n_segments/8) bytes, little endian */
    const uint8_t *units;        /* 4*n_units bytes */
    size_t total_bytes;          /* bytes of the image consumed */
} hd_huff_view;

hd_status hd_huff_parse(const void *file, size_t size, hd_huff_view *view);
size_t hd_huff_image_bytes(uint64_t symbol_count, uint32_t n_units,
                           uint32_t n_segments);

/* Whole-file convenience with HOST buffers (the `./bin/decoder in out` shape,
 * decoder/src/huff.cpp:31-129): parse, upload, decode on `device`, download.
 * out_cap must be >= original_size.  *out_size receives original_size.
 * timing_ms (may be NULL) receives {h2d, decode, d2h} like the reference's
 * "HtoD,dec,DtoH" line (decoder.cu:803). */
hd_status hd_huff_decode_host(const void *file, size_t size, uint8_t *out,
                              size_t out_cap, size_t *out_size, int device,
                              float timing_ms[3]);

/* Host-buffer decoder with persistent device buffers and its own streams: per
 * call = H2D + decode + D2H, the region the reference demos time
 * (gpuhd/src/demo.cc:170-197 "Decompression time"; decoder.cu:760-803
 * "HtoD,dec,DtoH"), without their per-call cudaMalloc.  The call is a
 * three-stage pipeline over pieces of the stream (2, 4, 8, then 16 MiB each):
 * uploads, decodes and downloads of different pieces overlap (PCIe in both
 * directions at once), so the wall time approaches max(H2D, D2H); the output
 * position of consecutive pieces is chained on the device, the host only
 * queues the downloads.  Host pointers should be pinned for full PCIe rate and
 * real overlap (the reference pins too, decoder/src/huff.cpp:44,88).
 * timing_ms (may be NULL) = host milliseconds {until the first piece is on the
 * device, from then until the last piece is decoded, waiting for the
 * remaining downloads}. */
typedef struct hd_host_decoder hd_host_decoder;
hd_status hd_host_decoder_create(int device, uint64_t max_units,
                                 uint64_t max_symbols, hd_host_decoder **out);
void hd_host_decoder_destroy(hd_host_decoder *h);
/* == CUHDGPUInputBuffer::cpy_host_to_device + CUHDGPUDecoder::decode +
 *    CUHDGPUOutputBuffer::cpy_device_to_host (gpuhd/src/demo.cc:183-195) */
hd_status hd_host_decode_selfsync(hd_host_decoder *h, const uint32_t *units,
                                  uint64_t n_units, const hd_lut_entry *lut,
                                  uint32_t lut_bits, uint8_t *out,
                                  uint64_t n_symbols, float timing_ms[3]);
/* == decoder_l1_l2 on a parsed .huff image (decoder/src/huff.cpp:81-118) */
hd_status hd_host_decode_huff(hd_host_decoder *h, const void *file,
                              size_t size, uint8_t *out, size_t out_cap,
                              size_t *out_size, float timing_ms[3]);
/* One rank's SHARD of a longer stream through the same pipeline (multi-GPU end
 * to end, see hd_mgpu_host_step).  units -> [left halo | shard | right halo] as
 * hd_mgpu_plan_shard lays it out; the gap variant takes the shard's own gap
 * elements and the offset at which its first segment starts.  info_out =
 * {symbols written to out, entry offset, exit offset}. */
hd_status hd_host_decode_selfsync_shard(hd_host_decoder *h, const uint32_t *units,
                                        uint64_t n_units_halo, uint64_t n_units_shard,
                                        uint64_t n_units_readable, const hd_table *table,
                                        uint8_t *out, uint64_t out_cap,
                                        uint64_t info_out[3], float timing_ms[3]);
hd_status hd_host_decode_gap_shard(hd_host_decoder *h, const uint32_t *units,
                                   uint64_t n_units_readable, const uint32_t *gap_words,
                                   uint64_t n_segments, uint32_t entry_bit,
                                   const hd_table *table, uint8_t *out, uint64_t out_cap,
                                   uint64_t info_out[3], float timing_ms[3]);
/* the decoder's device buffers (any pointer may be NULL) */
void hd_host_decoder_buffers(hd_host_decoder *h, uint32_t **d_units,
                             uint32_t **d_gap_words, uint8_t **d_out,
                             void **d_workspace, size_t *workspace_bytes);
/* one-shot `./bin/demo <gpu> <MB>` shape: allocate, H2D, decode, D2H, free */
hd_status hd_selfsync_decode_host(const uint32_t *units, uint64_t n_units,
                                  const hd_lut_entry *lut, uint32_t lut_bits,
                                  uint8_t *out, uint64_t n_symbols, int device,
                                  float timing_ms[3]);

/* --------------------------------------------------------------- profiling
 * Optional per-kernel device timing (the reference has cudaEvent TIMER_START/
 * TIMER_STOP around its kernel, decoder/include/cu_timer.h:10-37).  When
 * enabled on the calling thread, the next decode records one event pair per
 * kernel class on the decode stream; hd_profile_read returns the elapsed
 * milliseconds per slot and re-arms. */
#define HD_PROFILE_SYNC_CHUNKS 0   /* K1a */
#define HD_PROFILE_SYNC_SEAMS 1    /* K1b (both rounds) */
#define HD_PROFILE_DECODE_WRITE 2  /* K4 */
#define HD_PROFILE_SLOTS 4
/* Tuning knob for benchmarks: 128-bit segments decoded per thread (1, 2 or 4)
 * and decode threads per CTA (a multiple of 32, at most 864 = 27 warps) of the
 * decode-and-write kernel; 0, 0 = automatic (two segments per thread when at
 * least 20 decode warps fit into the SM's shared memory, else one).
 * segs_per_thread < 0 additionally keeps complete FIXED-LENGTH codes (every
 * codeword L bits, L in {1, 2, 4, 8}) on the general kernels; by default a
 * whole stream of such a code that starts at bit 0 and decodes into a 16-byte
 * aligned buffer is decoded by one dedicated kernel (no synchronisation, no
 * prescan: every segment starts at a codeword boundary and holds 128 / L
 * symbols). */
void hd_set_tuning(int segs_per_thread, int threads_per_tile);
/* The codeword length L if `table` is such a fixed-length code (the dedicated kernel is used for whole streams from
 * bit 0 into 16-byte aligned buffers), else 0. */
uint32_t hd_table_fixed_length(const hd_table *table);
/* Developer aid: device buffer (48 x 8 u64) that receives per-step %globaltimer
 * stamps of one CTA of the decode-and-write kernel; NULL switches it off. */
void hd_set_trace(void *d_buf);
/* kernels launched by this library in this process so far */
uint64_t hd_launch_count(void);
void hd_profile_enable(int on);
hd_status hd_profile_read(float ms_out[HD_PROFILE_SLOTS]);

/* ---------------------------------------------------- encoders (input side)
 * Host-side producers of both containers so that the CLIs and benchmarks can
 * make inputs without the reference tree.  NOT the hot path.
 * Length-limited (package-merge) code lengths + canonical codes:
 * gpuhd/encoder/src/llhuffman_encoder.cc:17-198 (limit 11),
 * Huffman_coding_Gap_arrays/encoder/src/package_merge.cpp:107-182 (limit 16).
 * pairs_out receives up to 256 pairs in non-decreasing length order. */
hd_status hd_build_code(const uint8_t *data, uint64_t n, uint32_t max_len,
                        hd_sym_len *pairs_out, uint32_t *n_pairs_out);
/* flat LUT of the code defined by pairs (canonical, list order) */
hd_status hd_flat_lut_from_lengths(const hd_sym_len *pairs, uint32_t n_pairs,
                                   uint32_t lut_bits, hd_lut_entry *lut_out);
/* number of units / segments the packed stream needs */
hd_status hd_encoded_size(const uint8_t *data, uint64_t n,
                          const hd_sym_len *pairs, uint32_t n_pairs,
                          uint64_t *bits_out);
/* multi-threaded MSB-first packer (always flushes the tail, unlike reference
 * defect D1); units_out: ceil(bits/32) units (+1 zero pad unit if
 * pad_unit != 0); gap_words_out (may be NULL): ceil(ceil(bits/128)/8) words */
hd_status hd_encode(const uint8_t *data, uint64_t n, const hd_sym_len *pairs,
                    uint32_t n_pairs, uint32_t *units_out, uint64_t cap_units,
                    int pad_unit, uint32_t *gap_words_out, int n_threads);
/* write a .huff image; returns bytes written through *written */
hd_status hd_huff_write(void *dst, size_t cap, const hd_sym_len *pairs,
                        uint32_t n_pairs, uint32_t original_size,
                        const uint32_t *units, uint32_t n_units,
                        const uint32_t *gap_words, uint32_t n_segments,
                        size_t *written);
/* synthetic data with the distribution of generate.cpp:33-47 (byte =
 * 'A'+U{0..3} with probability r, else U{0..255}); engine = mt19937_64 seeded
 * per 1 MiB block from (seed, block) so generation is parallel and
 * reproducible. */
hd_status hd_generate(uint8_t *out, uint64_t n, double redundancy,
                      uint64_t seed, int n_threads);

/* ------------------------------------------- device-side encoder (input side)
 * Producer of the .huff payload ON THE GPU, so that inputs of any size can be
 * made at device speed: byte histogram, then MSB-first bit packing with a
 * single-pass look-back over bit offsets, and the 4-bit gap array.  Replaces
 * cu_histgram/mergeHistogram, cuencoder and cu_get_gaparray
 * (Huffman_coding_Gap_arrays/encoder/src/encoder.cu:33-116, 142-355, 358-379);
 * the code itself is still built on the host from the 256-bin histogram
 * (boundary package-merge there, encoder/src/package_merge.cpp:107-182).
 * Output is bit-identical to hd_encode.  NOT the hot path.
 *  d_data 16-byte aligned; d_units/d_gap_words are zeroed by the call;
 *  cap_units >= ceil(bits/32) (+1 for the CUHD pad unit if wanted);
 *  *d_bits_out (device, may be NULL) receives the total number of code bits. */
size_t hd_workspace_bytes_encode(uint64_t n);
hd_status hd_histogram_device(const uint8_t *d_data, uint64_t n, uint64_t *d_hist256,
                              void *stream);
hd_status hd_build_code_from_histogram(const uint64_t *hist256, uint32_t max_len,
                                       hd_sym_len *pairs_out, uint32_t *n_pairs_out);
hd_status hd_encode_device(const uint8_t *d_data, uint64_t n, const hd_sym_len *pairs,
                           uint32_t n_pairs, uint32_t *d_units, uint64_t cap_units,
                           uint32_t *d_gap_words, uint64_t cap_gap_words,
                           void *d_workspace, size_t workspace_bytes,
                           uint64_t *d_bits_out, void *stream);

/* The same producer without the host in the loop: hd_build_code_device builds
 * the length-limited canonical code from a device-resident histogram with a
 * kernel (replaces the host package-merge, encoder/src/package_merge.cpp:
 * 107-182, and the header sort of encoder/src/huff.cpp:189-194; identical
 * result to hd_build_code_from_histogram) and also leaves the encoder's
 * 256-entry code table {code | len << 24} in d_code_tab256;
 * hd_encode_device_auto = histogram + code + pack + gap array on one stream,
 * nothing read back (d_pairs_out: 256 pairs, *d_n_pairs_out: how many; an
 * unusable histogram -- empty, or more symbols than 2^max_len -- gives 0).
 * Workspace of _auto: align64(hd_workspace_bytes_encode(n)) + 2048 +
 * hd_workspace_bytes_code(). */
size_t hd_workspace_bytes_code(void);
hd_status hd_build_code_device(const uint64_t *d_hist256, uint32_t max_len,
                               hd_sym_len *d_pairs_out, uint32_t *d_n_pairs_out,
                               uint32_t *d_code_tab256, void *d_scratch,
                               size_t scratch_bytes, void *stream);
hd_status hd_encode_device_auto(const uint8_t *d_data, uint64_t n, uint32_t max_len,
                                hd_sym_len *d_pairs_out, uint32_t *d_n_pairs_out,
                                uint32_t *d_units, uint64_t cap_units,
                                uint32_t *d_gap_words, uint64_t cap_gap_words,
                                void *d_workspace, size_t workspace_bytes,
                                uint64_t *d_bits_out, void *stream);

/* --------------------------------------------------------------- multi GPU
 * Sharding helpers (replace the naive split of gpuhd/multigpu_demo.cc:186-199).
 * Pure host arithmetic: shard `rank` of `world` over n_segments 128-bit
 * segments, aligned to whole tiles.  [seg_begin, seg_end). */
void hd_shard_segments(uint64_t n_segments, uint32_t rank, uint32_t world,
                       uint64_t *seg_begin, uint64_t *seg_end);

/* Mailbox: the per-step exchange of {count, entry, exit} between the ranks of
 * a sharded decode, done with peer-memory stores over NVLink / NVSwitch (CUDA
 * IPC) instead of a collective call: publish = one tiny kernel that writes
 * this rank's triple + a sequence number into its slot of EVERY peer's
 * mailbox; wait = one tiny kernel that returns once all slots of the local
 * mailbox carry that sequence number and copies the world x 3 values out.
 * Set-up (once): create on every rank, exchange the 64-byte handles by any
 * means (e.g. torch.distributed all_gather), connect.  seq must increase by
 * one per step and start at 1.  Replaces the host-side stitching of
 * gpuhd/multigpu_demo.cc:301-313; an ncclAllGather of the same 24 bytes is the
 * fallback (`sharding.py`). */
typedef struct hd_mailbox hd_mailbox;
hd_status hd_mailbox_create(uint32_t world, uint32_t rank, hd_mailbox **out);
hd_status hd_mailbox_export(hd_mailbox *m, uint8_t handle_out[64]);
hd_status hd_mailbox_connect(hd_mailbox *m, const uint8_t *handles /* world x 64 */);
hd_status hd_mailbox_publish(hd_mailbox *m, const uint64_t *d_info3, uint64_t seq,
                             void *stream);
hd_status hd_mailbox_wait(hd_mailbox *m, uint64_t seq, uint64_t *d_all_info3,
                          void *stream);
int hd_mailbox_timed_out(hd_mailbox *m);
void hd_mailbox_destroy(hd_mailbox *m);

/* One shard of a gap-array stream: like hd_decode_gap_ex, for a rank that holds
 * only its part of the stream.  n_units_readable counts the units from the
 * shard's first segment on that may be read (shard + right halo), n_segments
 * the segments to decode.  entry_from_prev_word != 0: the shard does not start
 * the stream; d_gap_words[-1] must be readable and its top nibble is the offset
 * at which the shard's first segment starts (decoder.cu:505-507).
 * d_shard_info receives {symbols written, entry offset, exit offset}. */
hd_status hd_decode_gap_shard(const uint32_t *d_units, uint64_t n_units_readable,
                              const uint32_t *d_gap_words, uint64_t n_segments,
                              int entry_from_prev_word, const hd_table *table,
                              uint8_t *d_out, uint64_t out_cap, void *d_workspace,
                              size_t workspace_bytes, uint64_t *d_shard_info,
                              void *stream);

/* hd_mgpu_*: the whole sharded decode of ONE stream over the GPUs of a box,
 * both schemes (replaces gpuhd/multigpu_demo.cc:178-313: partition :178-227,
 * per-GPU launch :255-298, host stitch :301-313).  One hd_mgpu per rank (one
 * process per GPU, or one host thread per GPU).  Life cycle:
 *   hd_mgpu_plan_shard   pure host arithmetic: which units / gap words this rank
 *                        uploads, buffer sizes
 *   hd_mgpu_create + one of hd_mgpu_exchange_{mailbox,nccl,callback}
 *   hd_mgpu_bind         device buffers of the rank (filled by the caller as the
 *                        plan says: d_units from plan.first_unit on,
 *                        d_gap_words from plan.first_gap_word on)
 *   hd_mgpu_enqueue      shard decode + exchange of {count, entry, exit} on the
 *                        stream, no host synchronisation (may be repeated)
 *   hd_mgpu_finish       COLLECTIVE host half: reads the exchanged triples,
 *                        repeats decodes whose synchronisation had not
 *                        converged, repairs seams (entry[g] != exit[g-1]: rank g
 *                        decodes again from the exact offset, a sequential
 *                        chain in the never-synchronising worst case), then
 *                        returns this rank's output placement: its bytes
 *                        d_out[0, count) belong at `offset` of the whole output.
 * No payload crosses NVLink; the output stays sharded. */
typedef enum { HD_SCHEME_SELFSYNC = 0, HD_SCHEME_GAP = 1 } hd_scheme;
typedef struct {
    uint64_t seg_begin, seg_end;   /* 128-bit segments [begin, end) of the stream decoded by this rank */
    uint64_t first_unit;           /* first unit of the stream the rank uploads */
    uint64_t halo_units;           /* self-sync: units before the shard, parsed speculatively */
    uint64_t shard_units;          /* units of the shard proper; 0 = nothing to decode (more ranks than tiles) */
    uint64_t readable_units;       /* units to upload from first_unit on: halo + shard + right halo */
    uint64_t first_gap_word;       /* gap scheme: first gap word the rank uploads ... */
    uint64_t n_gap_words;          /* ... and how many */
    uint64_t out_capacity;         /* bytes of the rank's output buffer (worst case of the code's shortest length) */
    uint64_t workspace_bytes;
} hd_mgpu_plan;
/* halo_segments: 0 = default (4096 segments = 64 KiB of stream) */
hd_status hd_mgpu_plan_shard(hd_scheme scheme, uint64_t n_units, uint32_t rank,
                             uint32_t world, uint32_t min_code_length,
                             uint64_t halo_segments, hd_mgpu_plan *plan);
/* pure host: the seam rule and the output placement over gathered triples
 * (world x {count, entry, exit}); first_bad_seam returns world if all agree */
uint32_t hd_mgpu_first_bad_seam(const uint64_t *infos, uint32_t world);
void hd_mgpu_place(const uint64_t *infos, uint32_t world, uint64_t n_symbols,
                   uint32_t rank, uint64_t *offset, uint64_t *count);

typedef struct hd_mgpu hd_mgpu;
/* all-gather of bytes_per_rank bytes per rank on `stream`; 0 = success */
typedef int (*hd_allgather_fn)(const void *d_send, void *d_recv,
                               size_t bytes_per_rank, void *stream, void *user);
hd_status hd_mgpu_create(hd_scheme scheme, uint32_t rank, uint32_t world,
                         const hd_mgpu_plan *plan, hd_mgpu **out);
void hd_mgpu_destroy(hd_mgpu *m);
/* the exchange: a connected mailbox handle -- peer memory over NVLink --, the rank's
 * ncclComm_t (passed as void*; ncclAllGather of the NCCL already loaded in the
 * process -- the library does not link against NCCL), or any callback */
hd_status hd_mgpu_exchange_mailbox(hd_mgpu *m, hd_mailbox *mailbox);
hd_status hd_mgpu_exchange_nccl(hd_mgpu *m, void *nccl_comm);
hd_status hd_mgpu_exchange_callback(hd_mgpu *m, hd_allgather_fn fn, void *user);
hd_status hd_mgpu_bind(hd_mgpu *m, const uint32_t *d_units,
                       const uint32_t *d_gap_words /* NULL for self-sync */,
                       const hd_table *table, uint8_t *d_out, void *d_workspace);
hd_status hd_mgpu_enqueue(hd_mgpu *m, void *stream);
/* one exchange round whose result nobody reads (benchmarks: rendezvous of the
 * ranks on the stream) */
hd_status hd_mgpu_align(hd_mgpu *m, void *stream);
hd_status hd_mgpu_finish(hd_mgpu *m, uint64_t n_symbols_total, uint64_t *offset,
                         uint64_t *count, void *stream);
/* End to end from / to HOST memory (the region the reference's multi-GPU demo
 * times, gpuhd/multigpu_demo.cc:255-313): the rank's shard runs through the
 * pipelined host decoder (H2D, decode and D2H of pieces overlap), then the
 * triple is exchanged like after hd_mgpu_enqueue; hd_mgpu_finish follows.  h_*
 * buffers should be pinned.  After a finish that had to repair this rank
 * (hd_mgpu_repair_count() != 0, never seen on self-synchronising codes) the
 * host output is refreshed with hd_mgpu_host_refresh. */
hd_status hd_mgpu_host_step(hd_mgpu *m, hd_host_decoder *hdec, const uint32_t *h_units,
                            const uint32_t *h_gap_words, uint8_t *h_out,
                            uint64_t out_cap, void *stream);
hd_status hd_mgpu_host_refresh(hd_mgpu *m, uint8_t *h_out, uint64_t count, void *stream);
/* the triples of the last hd_mgpu_finish (world x 3, host), and how many times
 * THIS rank had to decode again in it (0 in the normal case) */
hd_status hd_mgpu_infos(const hd_mgpu *m, uint64_t *infos_out);
uint32_t hd_mgpu_repair_count(const hd_mgpu *m);

#ifdef __cplusplus
}
#endif
#endif /* HD_B200_H_ */
