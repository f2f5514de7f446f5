"""B200-native parallel Huffman decoder -- Python host layer over the C ABI.

The product is ``libhd_b200.so`` (include/hd_b200.h): hand-written sm_100a CUDA
kernels behind a C ABI that mirrors the decode surface of
dek226/CSE375-FinalProj-Huffman-Decoding (gpuhd ``CUHDGPUDecoder::decode``,
gpuhd/include/cuhd_gpu_decoder.h:24-32, and the gap-array ``decoder_l1_l2`` /
``.huff`` file, Huffman_coding_Gap_arrays/decoder/include/decoder.cuh:4-15).
This module only moves pointers: torch supplies device memory, streams and
``torch.distributed``; every decode goes through the CUDA library and there is
no CPU fallback -- a missing library or device raises.

Because the package directory name contains hyphens (it is the reference's repo
name), import it with :func:`importlib` -- see ``hd_b200.py`` at the repo root.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HD_B200_LIB") or os.path.join(_HERE, "libhd_b200.so")   # override: developer builds only

HD_OK = 0
SEGMENT_UNITS = 4      # 128-bit subsequence / segment (gpuhd/src/demo.cc:16, encoder/include/constants.hpp:12)
TILE_SEGMENTS = 256
MAX_CODEWORD_LENGTH_CUHD = 11   # gpuhd/include/cuhd_constants.h:15
MAX_CODEWORD_LENGTH_GAP = 16    # Huffman_coding_Gap_arrays/decoder/include/constants.hpp:5


class HDError(RuntimeError):
    pass


class _HuffView(C.Structure):
    _fields_ = [("symbol_count", C.c_uint64), ("symbols", C.c_void_p), ("original_size", C.c_uint32),
                ("n_units", C.c_uint32), ("n_segments", C.c_uint32), ("gap_words", C.c_void_p),
                ("units", C.c_void_p), ("total_bytes", C.c_size_t)]


_lib = None

# every symbol include/hd_b200.h declares (tests check that all are exported)
ABI_SYMBOLS = [
    "hd_version", "hd_status_string", "hd_last_cuda_error", "hd_warmup", "hd_table_from_flat_lut", "hd_table_from_lengths",
    "hd_table_free", "hd_table_max_length", "hd_table_min_length", "hd_table_fixed_length", "hd_workspace_bytes_gap",
    "hd_workspace_bytes_selfsync", "hd_decode_gap", "hd_decode_selfsync", "hd_decode_gap_ex",
    "hd_decode_selfsync_ex", "hd_selfsync_gap_array", "hd_huff_parse", "hd_huff_image_bytes",
    "hd_huff_decode_host", "hd_build_code", "hd_flat_lut_from_lengths", "hd_encoded_size", "hd_encode",
    "hd_huff_write", "hd_generate", "hd_shard_segments", "hd_host_decoder_create", "hd_host_decoder_destroy",
    "hd_host_decode_selfsync", "hd_host_decode_huff", "hd_selfsync_decode_host", "hd_profile_enable",
    "hd_profile_read", "hd_launch_count", "hd_decode_selfsync_shard", "hd_set_tuning", "hd_set_trace", "hd_decode_selfsync_enqueue",
    "hd_decode_selfsync_finish", "hd_table_from_lengths_device", "hd_table_blob",
    "hd_decode_selfsync_shard_enqueue", "hd_decode_selfsync_shard_finish", "hd_workspace_bytes_encode",
    "hd_histogram_device", "hd_build_code_from_histogram", "hd_encode_device", "hd_mailbox_create",
    "hd_mailbox_export", "hd_mailbox_connect", "hd_mailbox_publish", "hd_mailbox_wait", "hd_mailbox_timed_out",
    "hd_mailbox_destroy", "hd_decode_gap_shard", "hd_mgpu_plan_shard", "hd_mgpu_first_bad_seam", "hd_mgpu_place",
    "hd_mgpu_create", "hd_mgpu_destroy", "hd_mgpu_exchange_mailbox", "hd_mgpu_exchange_nccl",
    "hd_mgpu_exchange_callback", "hd_mgpu_bind", "hd_mgpu_enqueue", "hd_mgpu_align", "hd_mgpu_finish", "hd_mgpu_infos",
    "hd_workspace_bytes_code", "hd_build_code_device", "hd_encode_device_auto",
    "hd_mgpu_repair_count", "hd_mgpu_host_step", "hd_mgpu_host_refresh", "hd_host_decode_selfsync_shard",
    "hd_host_decode_gap_shard", "hd_host_decoder_buffers",
]
PROFILE_SLOTS = ("sync_chunks", "sync_seams", "decode_write", "reserved")


def lib():
    """Load the CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HDError("libhd_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(or `make -C cse375-finalproj-huffman-decoding_b200`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
    L.hd_version.restype = C.c_char_p
    L.hd_status_string.restype = C.c_char_p
    L.hd_status_string.argtypes = [C.c_int]
    L.hd_last_cuda_error.restype = C.c_char_p
    L.hd_table_from_flat_lut.argtypes = [vp, u32, vp, C.POINTER(vp)]
    L.hd_table_from_lengths.argtypes = [vp, u32, vp, C.POINTER(vp)]
    L.hd_table_free.argtypes = [vp]
    L.hd_table_from_lengths_device.argtypes = [vp, u32, vp, C.POINTER(vp)]
    L.hd_table_blob.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.hd_table_free.restype = None
    L.hd_table_max_length.argtypes = [vp]
    L.hd_table_max_length.restype = u32
    L.hd_table_min_length.argtypes = [vp]
    L.hd_table_min_length.restype = u32
    L.hd_table_fixed_length.argtypes = [vp]
    L.hd_table_fixed_length.restype = u32
    L.hd_workspace_bytes_gap.argtypes = [u64]
    L.hd_workspace_bytes_gap.restype = sz
    L.hd_workspace_bytes_selfsync.argtypes = [u64]
    L.hd_workspace_bytes_selfsync.restype = sz
    L.hd_decode_gap.argtypes = [vp, u64, vp, u64, vp, vp, u64, vp, sz, vp]
    L.hd_decode_gap_ex.argtypes = [vp, u64, vp, u64, u32, vp, vp, u64, vp, sz, vp, vp]
    L.hd_decode_selfsync.argtypes = [vp, u64, vp, vp, u64, vp, sz, vp]
    L.hd_decode_selfsync_ex.argtypes = [vp, u64, u32, vp, vp, u64, vp, sz, vp, vp]
    L.hd_selfsync_gap_array.argtypes = [vp, u64, u32, vp, vp, vp, sz, vp]
    L.hd_decode_selfsync_enqueue.argtypes = [vp, u64, u32, vp, vp, u64, vp, sz, vp, vp, vp]
    L.hd_decode_selfsync_finish.argtypes = [vp, u64, u32, vp, vp, u64, vp, sz, vp, vp, vp]
    L.hd_huff_parse.argtypes = [vp, sz, C.POINTER(_HuffView)]
    L.hd_huff_image_bytes.argtypes = [u64, u32, u32]
    L.hd_huff_image_bytes.restype = sz
    L.hd_huff_decode_host.argtypes = [vp, sz, vp, sz, C.POINTER(sz), C.c_int, vp]
    L.hd_build_code.argtypes = [vp, u64, u32, vp, C.POINTER(u32)]
    L.hd_flat_lut_from_lengths.argtypes = [vp, u32, u32, vp]
    L.hd_encoded_size.argtypes = [vp, u64, vp, u32, C.POINTER(u64)]
    L.hd_encode.argtypes = [vp, u64, vp, u32, vp, u64, C.c_int, vp, C.c_int]
    L.hd_huff_write.argtypes = [vp, sz, vp, u32, u32, vp, u32, vp, u32, C.POINTER(sz)]
    L.hd_generate.argtypes = [vp, u64, C.c_double, u64, C.c_int]
    L.hd_shard_segments.argtypes = [u64, u32, u32, C.POINTER(u64), C.POINTER(u64)]
    L.hd_shard_segments.restype = None
    L.hd_host_decoder_create.argtypes = [C.c_int, u64, u64, C.POINTER(vp)]
    L.hd_host_decoder_destroy.argtypes = [vp]
    L.hd_host_decoder_destroy.restype = None
    L.hd_host_decode_selfsync.argtypes = [vp, vp, u64, vp, u32, vp, u64, vp]
    L.hd_host_decode_huff.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), vp]
    L.hd_selfsync_decode_host.argtypes = [vp, u64, vp, u32, vp, u64, C.c_int, vp]
    L.hd_profile_enable.argtypes = [C.c_int]
    L.hd_profile_enable.restype = None
    L.hd_profile_read.argtypes = [vp]
    L.hd_launch_count.restype = u64
    L.hd_mailbox_create.argtypes = [u32, u32, C.POINTER(vp)]
    L.hd_mailbox_export.argtypes = [vp, vp]
    L.hd_mailbox_connect.argtypes = [vp, vp]
    L.hd_mailbox_publish.argtypes = [vp, vp, u64, vp]
    L.hd_mailbox_wait.argtypes = [vp, u64, vp, vp]
    L.hd_mailbox_timed_out.argtypes = [vp]
    L.hd_mailbox_timed_out.restype = C.c_int
    L.hd_mailbox_destroy.argtypes = [vp]
    L.hd_mailbox_destroy.restype = None
    L.hd_workspace_bytes_encode.argtypes = [u64]
    L.hd_workspace_bytes_encode.restype = sz
    L.hd_histogram_device.argtypes = [vp, u64, vp, vp]
    L.hd_build_code_from_histogram.argtypes = [vp, u32, vp, C.POINTER(u32)]
    L.hd_encode_device.argtypes = [vp, u64, vp, u32, vp, u64, vp, u64, vp, sz, vp, vp]
    L.hd_workspace_bytes_code.restype = sz
    L.hd_build_code_device.argtypes = [vp, u32, vp, vp, vp, vp, sz, vp]
    L.hd_encode_device_auto.argtypes = [vp, u64, u32, vp, vp, vp, u64, vp, u64, vp, sz, vp, vp]
    L.hd_set_tuning.argtypes = [C.c_int, C.c_int]
    L.hd_set_tuning.restype = None
    L.hd_set_trace.argtypes = [vp]
    L.hd_set_trace.restype = None
    L.hd_decode_selfsync_shard.argtypes = [vp, u64, u64, u64, u32, vp, vp, u64, vp, sz, vp, vp]
    L.hd_decode_selfsync_shard_enqueue.argtypes = [vp, u64, u64, u64, u32, vp, vp, u64, vp, sz, vp, vp, vp]
    L.hd_decode_selfsync_shard_finish.argtypes = [vp, u64, u64, u64, u32, vp, vp, u64, vp, sz, vp, vp, vp]
    L.hd_decode_gap_shard.argtypes = [vp, u64, vp, u64, C.c_int, vp, vp, u64, vp, sz, vp, vp]
    L.hd_mgpu_plan_shard.argtypes = [C.c_int, u64, u32, u32, u32, u64, vp]
    L.hd_mgpu_first_bad_seam.argtypes = [vp, u32]
    L.hd_mgpu_first_bad_seam.restype = u32
    L.hd_mgpu_place.argtypes = [vp, u32, u64, u32, C.POINTER(u64), C.POINTER(u64)]
    L.hd_mgpu_place.restype = None
    L.hd_mgpu_create.argtypes = [C.c_int, u32, u32, vp, C.POINTER(vp)]
    L.hd_mgpu_destroy.argtypes = [vp]
    L.hd_mgpu_destroy.restype = None
    L.hd_mgpu_exchange_mailbox.argtypes = [vp, vp]
    L.hd_mgpu_exchange_nccl.argtypes = [vp, vp]
    L.hd_mgpu_exchange_callback.argtypes = [vp, vp, vp]
    L.hd_mgpu_bind.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hd_mgpu_enqueue.argtypes = [vp, vp]
    L.hd_mgpu_align.argtypes = [vp, vp]
    L.hd_mgpu_finish.argtypes = [vp, u64, C.POINTER(u64), C.POINTER(u64), vp]
    L.hd_mgpu_infos.argtypes = [vp, vp]
    L.hd_mgpu_repair_count.argtypes = [vp]
    L.hd_mgpu_host_step.argtypes = [vp, vp, vp, vp, vp, u64, vp]
    L.hd_mgpu_host_refresh.argtypes = [vp, vp, u64, vp]
    L.hd_host_decode_selfsync_shard.argtypes = [vp, vp, u64, u64, u64, vp, vp, u64, vp, vp]
    L.hd_host_decode_gap_shard.argtypes = [vp, vp, u64, vp, u64, u32, vp, vp, u64, vp, vp]
    L.hd_host_decoder_buffers.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hd_host_decoder_buffers.restype = None
    L.hd_mgpu_repair_count.restype = u32
    _lib = L
    return L


def _check(st, what):
    if st != HD_OK:
        L = lib()
        msg = L.hd_status_string(st).decode()
        if st == 2:
            msg += ": " + L.hd_last_cuda_error().decode()
        raise HDError("%s failed: %s" % (what, msg))


def version():
    return lib().hd_version().decode()


# --------------------------------------------------------------------------
# host side: code construction, packing, containers, synthetic data
# --------------------------------------------------------------------------
def generate(n, redundancy, seed=1, threads=0):
    """Synthetic bytes with the distribution of generate.cpp:33-47."""
    out = np.empty(n, dtype=np.uint8)
    _check(lib().hd_generate(out.ctypes.data, n, float(redundancy), int(seed), threads), "hd_generate")
    return out


def build_code(data, max_len):
    """Length-limited canonical code -> (pairs[n,2] uint8 {symbol,length}, non-decreasing length)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    pairs = np.zeros((256, 2), dtype=np.uint8)
    n = C.c_uint32(0)
    _check(lib().hd_build_code(data.ctypes.data, len(data), max_len, pairs.ctypes.data, C.byref(n)), "hd_build_code")
    return pairs[:n.value].copy()


def flat_lut(pairs, lut_bits):
    """The reference's flat LUT (2^lut_bits x {num_bits,symbol}) of a canonical code."""
    pairs = np.ascontiguousarray(pairs, dtype=np.uint8)
    lut = np.zeros((1 << lut_bits, 2), dtype=np.uint8)
    _check(lib().hd_flat_lut_from_lengths(pairs.ctypes.data, len(pairs), lut_bits, lut.ctypes.data),
           "hd_flat_lut_from_lengths")
    return lut


def encode(data, pairs, pad_unit=False, want_gaps=False, threads=0):
    """MSB-first packing -> (units uint32[, gap_words uint32], total_bits)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint8)
    bits = C.c_uint64(0)
    _check(lib().hd_encoded_size(data.ctypes.data, len(data), pairs.ctypes.data, len(pairs), C.byref(bits)),
           "hd_encoded_size")
    n_units = (bits.value + 31) // 32
    n_seg = (bits.value + 127) // 128
    units = np.zeros(n_units + (1 if pad_unit else 0), dtype=np.uint32)
    gaps = np.zeros((n_seg + 7) // 8, dtype=np.uint32) if want_gaps else None
    _check(lib().hd_encode(data.ctypes.data, len(data), pairs.ctypes.data, len(pairs), units.ctypes.data,
                           len(units), 1 if pad_unit else 0, gaps.ctypes.data if want_gaps else None, threads),
           "hd_encode")
    if want_gaps:
        return units, gaps, bits.value
    return units, bits.value


def huff_write(pairs, original_size, units, gap_words):
    """Serialise a .huff image (layout of encoder/src/huff.cpp:186-202)."""
    pairs = np.ascontiguousarray(pairs, dtype=np.uint8)
    units = np.ascontiguousarray(units, dtype=np.uint32)
    gap_words = np.ascontiguousarray(gap_words, dtype=np.uint32)
    n_seg = (len(units) + 3) // 4
    size = lib().hd_huff_image_bytes(len(pairs), len(units), n_seg)
    buf = np.zeros(size, dtype=np.uint8)
    w = C.c_size_t(0)
    _check(lib().hd_huff_write(buf.ctypes.data, size, pairs.ctypes.data, len(pairs), original_size,
                               units.ctypes.data, len(units), gap_words.ctypes.data, n_seg, C.byref(w)),
           "hd_huff_write")
    return buf[:w.value]


class HuffFile:
    """Parsed view of a .huff image (reader of decoder/src/huff.cpp:36-100)."""

    def __init__(self, image):
        self.image = np.frombuffer(bytes(image), dtype=np.uint8) if not isinstance(image, np.ndarray) \
            else np.ascontiguousarray(image, dtype=np.uint8)
        v = _HuffView()
        _check(lib().hd_huff_parse(self.image.ctypes.data, len(self.image), C.byref(v)), "hd_huff_parse")
        base = self.image.ctypes.data
        self.symbol_count = int(v.symbol_count)
        self.original_size = int(v.original_size)
        self.n_units = int(v.n_units)
        self.n_segments = int(v.n_segments)
        so = v.symbols - base
        self.pairs = self.image[so:so + 2 * self.symbol_count].reshape(-1, 2).copy()
        go = v.gap_words - base
        n_gw = (self.n_segments + 7) // 8
        self.gap_words = self.image[go:go + 4 * n_gw].copy().view(np.uint32)
        uo = v.units - base
        self.units = self.image[uo:uo + 4 * self.n_units].copy().view(np.uint32)


def shard_segments(n_segments, rank, world):
    b, e = C.c_uint64(0), C.c_uint64(0)
    lib().hd_shard_segments(n_segments, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


# --------------------------------------------------------------------------
# device side (torch = memory + streams only)
# --------------------------------------------------------------------------
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise HDError("no CUDA device: the decoder has no CPU fallback")
    return torch


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Codetable:
    """Device decode tables; mirrors cuhd::CUHDGPUCodetable (gpuhd/include/cuhd_gpu_codetable.h)."""

    def __init__(self, flat_lut=None, pairs=None, device_pairs=None):
        torch = _torch()
        h = C.c_void_p(None)
        if device_pairs is not None:     # uint8 CUDA tensor [n, 2] {symbol, length}: tables are built by a kernel
            _check(lib().hd_table_from_lengths_device(device_pairs.data_ptr(), device_pairs.shape[0],
                                                      _stream_ptr(torch), C.byref(h)), "hd_table_from_lengths_device")
        elif flat_lut is not None:
            f = np.ascontiguousarray(flat_lut, dtype=np.uint8)
            bits = int(f.shape[0]).bit_length() - 1
            _check(lib().hd_table_from_flat_lut(f.ctypes.data, bits, _stream_ptr(torch), C.byref(h)),
                   "hd_table_from_flat_lut")
        else:
            p = np.ascontiguousarray(pairs, dtype=np.uint8)
            _check(lib().hd_table_from_lengths(p.ctypes.data, len(p), _stream_ptr(torch), C.byref(h)),
                   "hd_table_from_lengths")
        torch.cuda.current_stream().synchronize()   # the host blob may go away after this
        self.handle = h

    def blob(self):
        """The device table blob as a numpy byte array (for tests)."""
        size = C.c_size_t(0)
        _check(lib().hd_table_blob(self.handle, None, 0, C.byref(size)), "hd_table_blob")
        out = np.zeros(size.value, dtype=np.uint8)
        _check(lib().hd_table_blob(self.handle, out.ctypes.data, size.value, C.byref(size)), "hd_table_blob")
        return out

    @property
    def max_length(self):
        return lib().hd_table_max_length(self.handle)

    @property
    def min_length(self):
        return lib().hd_table_min_length(self.handle)

    def fixed_length(self):
        """L if every codeword is L bits, L in {1, 2, 4, 8} (such streams take the dedicated kernel), else 0."""
        return int(lib().hd_table_fixed_length(self.handle))

    def close(self):
        if self.handle:
            lib().hd_table_free(self.handle)
            self.handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def to_device_units(units):
    """uint32 units -> int32 CUDA tensor with a 16-byte aligned base."""
    torch = _torch()
    u = np.ascontiguousarray(units, dtype=np.uint32)
    return torch.from_numpy(u.view(np.int32)).cuda()


class Decoder:
    """Reusable workspace + output buffer; mirrors cuhd::CUHDGPUDecoderMemory + CUHDGPUDecoder."""

    def __init__(self, n_units, selfsync=True):
        torch = _torch()
        self.n_units = int(n_units)
        nbytes = lib().hd_workspace_bytes_selfsync(self.n_units) if selfsync else lib().hd_workspace_bytes_gap(self.n_units)
        self.ws = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
        off = (-self.ws.data_ptr()) % 64
        self.ws_ptr = self.ws.data_ptr() + off
        self.ws_bytes = nbytes
        self.count = torch.zeros(1, dtype=torch.int64, device="cuda")

    def decode_selfsync(self, d_units, table, n_symbols, out=None, first_bit=0, defer=False):
        """cuhd::CUHDGPUDecoder::decode (gpuhd/src/cuhd_gpu_decoder.cu:422-523).  defer=True only enqueues the work
        (hd_decode_selfsync_enqueue); call finish() before using the output."""
        torch = _torch()
        if out is None:
            out = torch.empty(n_symbols, dtype=torch.uint8, device="cuda")
        if not defer:
            _check(lib().hd_decode_selfsync_ex(d_units.data_ptr(), d_units.numel(), first_bit, table.handle,
                                               out.data_ptr(), n_symbols, self.ws_ptr, self.ws_bytes,
                                               self.count.data_ptr(), _stream_ptr(torch)), "hd_decode_selfsync")
            return out
        if getattr(self, "_flag", None) is None:
            self._flag = torch.zeros(16, dtype=torch.int32).pin_memory()
        self._pending = (d_units.data_ptr(), d_units.numel(), first_bit, table.handle, out.data_ptr(), n_symbols,
                         self.ws_ptr, self.ws_bytes, self.count.data_ptr(), self._flag.data_ptr(), _stream_ptr(torch))
        self._keep = (d_units, table, out)
        _check(lib().hd_decode_selfsync_enqueue(*self._pending), "hd_decode_selfsync_enqueue")
        return out

    def finish(self):
        """Second half of a deferred decode_selfsync: synchronise, check convergence, repair if needed."""
        if getattr(self, "_pending", None) is not None:
            _check(lib().hd_decode_selfsync_finish(*self._pending), "hd_decode_selfsync_finish")
            self._pending = None

    def decode_gap(self, d_units, d_gap_words, table, n_symbols, out=None, first_bit=0, n_segments=None):
        """decoder_l1_l2 / gpu_dec_l1_l2 (Huffman_coding_Gap_arrays/decoder/src/decoder.cu:454-815)."""
        torch = _torch()
        if out is None:
            out = torch.empty(n_symbols, dtype=torch.uint8, device="cuda")
        n_seg = (d_units.numel() + 3) // 4 if n_segments is None else int(n_segments)
        _check(lib().hd_decode_gap_ex(d_units.data_ptr(), d_units.numel(),
                                      d_gap_words.data_ptr() if d_gap_words is not None else None, n_seg,
                                      first_bit, table.handle, out.data_ptr(), n_symbols, self.ws_ptr,
                                      self.ws_bytes, self.count.data_ptr(), _stream_ptr(torch)), "hd_decode_gap")
        return out

    def gap_array(self, d_units, table, first_bit=0):
        """Phases 1+2 only: the stream's gap array (cuhd_gpu_decoder.cu:145-327 fixed point)."""
        torch = _torch()
        n_seg = (d_units.numel() + 3) // 4
        gw = torch.zeros((n_seg + 7) // 8, dtype=torch.int32, device="cuda")
        _check(lib().hd_selfsync_gap_array(d_units.data_ptr(), d_units.numel(), first_bit, table.handle,
                                           gw.data_ptr(), self.ws_ptr, self.ws_bytes, _stream_ptr(torch)),
               "hd_selfsync_gap_array")
        return gw


def decode_selfsync(units, flat_lut, n_symbols):
    """One-shot helper with HOST arrays (the `./bin/demo` shape, gpuhd/src/demo.cc:155-197)."""
    torch = _torch()
    d_units = to_device_units(units)
    tab = Codetable(flat_lut=flat_lut)
    dec = Decoder(d_units.numel(), selfsync=True)
    out = dec.decode_selfsync(d_units, tab, n_symbols)
    res = out.cpu().numpy()
    tab.close()
    return res


def decode_huff(image, device=-1):
    """One-shot `.huff` image -> bytes through hd_huff_decode_host (`./bin/decoder in out` shape)."""
    _torch()
    img = np.frombuffer(bytes(image), dtype=np.uint8) if not isinstance(image, np.ndarray) \
        else np.ascontiguousarray(image, dtype=np.uint8)
    hf = HuffFile(img)
    out = np.zeros(max(hf.original_size, 1), dtype=np.uint8)
    osz = C.c_size_t(0)
    t = (C.c_float * 3)()
    _check(lib().hd_huff_decode_host(img.ctypes.data, len(img), out.ctypes.data, len(out), C.byref(osz), device, t),
           "hd_huff_decode_host")
    return out[:osz.value], [t[0], t[1], t[2]]


class HostDecoder:
    """Host-buffer decoder (H2D + decode + D2H per call) with persistent device buffers: the region the reference
    demos time (gpuhd/src/demo.cc:170-197; decoder.cu:760-803).  Inputs/outputs are numpy arrays or pinned torch
    CPU tensors; everything goes through hd_host_decode_* of the C ABI."""

    def __init__(self, max_units, max_symbols, device=0):
        _torch()
        h = C.c_void_p(None)
        _check(lib().hd_host_decoder_create(device, int(max_units), int(max_symbols), C.byref(h)),
               "hd_host_decoder_create")
        self.handle = h

    @staticmethod
    def _ptr(a):
        return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

    def decode_selfsync(self, units, flat_lut, out, n_symbols):
        f = np.ascontiguousarray(flat_lut, dtype=np.uint8)
        bits = int(f.shape[0]).bit_length() - 1
        t = (C.c_float * 3)()
        n_units = units.numel() if hasattr(units, "numel") else len(units)
        _check(lib().hd_host_decode_selfsync(self.handle, self._ptr(units), n_units, f.ctypes.data, bits,
                                             self._ptr(out), int(n_symbols), t), "hd_host_decode_selfsync")
        return [t[0], t[1], t[2]]

    def decode_huff(self, image, out):
        t = (C.c_float * 3)()
        osz = C.c_size_t(0)
        n = image.numel() if hasattr(image, "numel") else len(image)
        cap = out.numel() if hasattr(out, "numel") else len(out)
        _check(lib().hd_host_decode_huff(self.handle, self._ptr(image), n, self._ptr(out), cap, C.byref(osz), t),
               "hd_host_decode_huff")
        return osz.value, [t[0], t[1], t[2]]

    def close(self):
        if self.handle:
            lib().hd_host_decoder_destroy(self.handle)
            self.handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def profile_enable(on=True):
    lib().hd_profile_enable(1 if on else 0)


def profile_read():
    """{slot: ms} of the kernels of the last profiled decode (hd_profile_read)."""
    t = (C.c_float * 4)()
    _check(lib().hd_profile_read(t), "hd_profile_read")
    return dict(zip(PROFILE_SLOTS, [t[0], t[1], t[2], t[3]]))


def set_tuning(segs_per_thread=0, threads_per_tile=0):
    lib().hd_set_tuning(int(segs_per_thread), int(threads_per_tile))


def launch_count():
    return int(lib().hd_launch_count())


from .sharding import (ShardedSelfSyncDecoder, ShardedGapDecoder, plan_shard, plan_shards, place_outputs, seams_ok,  # noqa: E402
                       resolve_seams, SCHEME_SELFSYNC, SCHEME_GAP)


def encode_device(d_data, max_len, pad_unit=False):
    """GPU encoder (hd_histogram_device + host package-merge + hd_encode_device): uint8 CUDA tensor ->
    (pairs, d_units int32 tensor, d_gap_words int32 tensor, total_bits).  Input side, not the hot path."""
    torch = _torch()
    n = d_data.numel()
    hist = torch.zeros(256, dtype=torch.int64, device="cuda")
    _check(lib().hd_histogram_device(d_data.data_ptr(), n, hist.data_ptr(), _stream_ptr(torch)), "hd_histogram_device")
    h = hist.cpu().numpy().astype(np.uint64)
    pairs = np.zeros((256, 2), dtype=np.uint8)
    k = C.c_uint32(0)
    _check(lib().hd_build_code_from_histogram(h.ctypes.data, max_len, pairs.ctypes.data, C.byref(k)),
           "hd_build_code_from_histogram")
    pairs = pairs[:k.value].copy()
    lens = np.zeros(256, dtype=np.uint64)
    lens[pairs[:, 0]] = pairs[:, 1]
    bits = int((h * lens).sum())
    n_units = (bits + 31) // 32
    n_seg = (bits + 127) // 128
    d_units = torch.empty(n_units + (1 if pad_unit else 0) + 4, dtype=torch.int32, device="cuda")
    d_gaps = torch.empty((n_seg + 7) // 8 + 1, dtype=torch.int32, device="cuda")
    nbytes = lib().hd_workspace_bytes_encode(n)
    ws = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 64
    d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
    _check(lib().hd_encode_device(d_data.data_ptr(), n, pairs.ctypes.data, len(pairs), d_units.data_ptr(),
                                  d_units.numel(), d_gaps.data_ptr(), d_gaps.numel(), ws_ptr, nbytes, d_bits.data_ptr(),
                                  _stream_ptr(torch)), "hd_encode_device")
    torch.cuda.current_stream().synchronize()
    assert int(d_bits.item()) == bits
    return pairs, d_units[:n_units + (1 if pad_unit else 0)], d_gaps[:(n_seg + 7) // 8], bits


def encode_device_auto(d_data, max_len, pad_unit=False):
    """The whole producer on the GPU without the host in the loop (hd_encode_device_auto: histogram -> length-limited
    canonical code built by a kernel -> pack + gap array); only the 256-pair header list and the bit count are read back
    afterwards.  Returns what encode_device returns."""
    torch = _torch()
    n = d_data.numel()
    cap_units = n * 9 // 32 + 16 if max_len >= 9 else n * max_len // 32 + 16      # a length-limited Huffman code averages < 9 bits
    d_units = torch.empty(cap_units, dtype=torch.int32, device="cuda")
    d_gaps = torch.empty((cap_units // 4 + 8) // 8 + 8, dtype=torch.int32, device="cuda")
    enc_bytes = (lib().hd_workspace_bytes_encode(n) + 63) // 64 * 64
    nbytes = enc_bytes + 2048 + lib().hd_workspace_bytes_code()
    ws = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 64
    d_pairs = torch.zeros(512, dtype=torch.uint8, device="cuda")
    d_meta = torch.zeros(2, dtype=torch.int64, device="cuda")      # [n_pairs (u32), total bits]
    _check(lib().hd_encode_device_auto(d_data.data_ptr(), n, max_len, d_pairs.data_ptr(), d_meta.data_ptr(),
                                       d_units.data_ptr(), cap_units, d_gaps.data_ptr(), d_gaps.numel(), ws_ptr, nbytes,
                                       d_meta.data_ptr() + 8, _stream_ptr(torch)), "hd_encode_device_auto")
    meta = d_meta.cpu().numpy()
    k, bits = int(meta[0]) & 0xffffffff, int(meta[1])
    if k == 0:
        raise HDError("hd_encode_device_auto: no code for this input (empty, or more symbols than 2^max_len)")
    pairs = d_pairs.cpu().numpy()[:2 * k].reshape(-1, 2).copy()
    n_units = (bits + 31) // 32
    n_seg = (bits + 127) // 128
    return pairs, d_units[:n_units + (1 if pad_unit else 0)], d_gaps[:(n_seg + 7) // 8], bits
