"""ctypes bindings for the CPU oracle (oracle/libhd_oracle.so) and, when present,
the reference-built checkers in oracle/_ref/.  Test infrastructure only -- the
product package never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libhd_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("hd_oracle.c", "hd_oracle.h", "gen_restatement.cpp")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libhd_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.oracle_canonical_codes.restype = C.c_int
        L.oracle_canonical_codes.argtypes = [_u8p, _u8p, C.c_int, _u32p, _u8p]
        L.oracle_build_flat_lut.restype = None
        L.oracle_build_flat_lut.argtypes = [_u32p, _u8p, C.c_uint32, C.c_void_p]
        L.oracle_pack.restype = C.c_uint64
        L.oracle_pack.argtypes = [_u8p, C.c_uint64, _u32p, _u8p, _u32p, C.c_uint64, C.c_int, _u8p, C.c_uint64]
        L.oracle_pack_gap_words.restype = None
        L.oracle_pack_gap_words.argtypes = [_u8p, C.c_uint64, _u32p]
        L.oracle_decode_flat.restype = C.c_uint64
        L.oracle_decode_flat.argtypes = [_u32p, C.c_uint64, C.c_void_p, C.c_uint32, _u8p, C.c_uint64]
        L.oracle_sync_points.restype = C.c_uint64
        L.oracle_sync_points.argtypes = [_u32p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32,
                                         _u8p, _u32p, C.c_uint64]
        L.oracle_decode_gap.restype = C.c_uint64
        L.oracle_decode_gap.argtypes = [_u32p, C.c_uint64, _u32p, C.c_uint64, C.c_void_p, C.c_uint32, _u8p,
                                        C.c_uint64]
        L.oracle_huff_size.restype = C.c_size_t
        L.oracle_huff_size.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.oracle_huff_write.restype = C.c_size_t
        L.oracle_huff_write.argtypes = [C.c_void_p, _u8p, _u8p, C.c_uint64, C.c_uint32, _u32p, C.c_uint32, _u32p,
                                        C.c_uint32]
        L.oracle_huff_decode.restype = C.c_int64
        L.oracle_huff_decode.argtypes = [C.c_void_p, C.c_size_t, _u8p, C.c_uint64]
        L.oracle_generate.restype = None
        L.oracle_generate.argtypes = [_u8p, C.c_uint64, C.c_double, C.c_uint64]
        _lib = L
    return _lib


# ---------------------------------------------------------------- oracle API
def generate(n, redundancy, seed):
    """Seeded restatement of generate.cpp:32-47."""
    out = np.empty(n, dtype=np.uint8)
    if n:
        lib().oracle_generate(_ptr(out, _u8p), n, float(redundancy), int(seed))
    return out


def canonical_codes(syms, lengths):
    syms = np.ascontiguousarray(syms, dtype=np.uint8)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
    codes = np.zeros(256, dtype=np.uint32)
    lens = np.zeros(256, dtype=np.uint8)
    rc = lib().oracle_canonical_codes(_ptr(syms, _u8p), _ptr(lengths, _u8p), len(syms), _ptr(codes, _u32p),
                                      _ptr(lens, _u8p))
    if rc:
        raise ValueError("bad (symbol,length) list")
    return codes, lens


def flat_lut(codes, lens, lut_bits):
    lut = np.zeros((1 << lut_bits, 2), dtype=np.uint8)
    lib().oracle_build_flat_lut(_ptr(codes, _u32p), _ptr(lens, _u8p), lut_bits, lut.ctypes.data)
    return lut


def pack(data, codes, lens, flush_tail=True, want_gaps=False):
    """Returns (units, total_bits[, gap_nibbles])."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    bits = int(lens[data].astype(np.uint64).sum()) if len(data) else 0
    n_units = (bits + 31) // 32
    n_seg = (bits + 127) // 128
    units = np.zeros(max(n_units, 1), dtype=np.uint32)
    gaps = np.zeros(max(n_seg, 1), dtype=np.uint8)
    got = lib().oracle_pack(_ptr(data, _u8p), len(data), _ptr(codes, _u32p), _ptr(lens, _u8p), _ptr(units, _u32p),
                            n_units, 1 if flush_tail else 0, _ptr(gaps, _u8p) if want_gaps else None, n_seg)
    assert got == bits
    units = units[:n_units]
    if want_gaps:
        return units, bits, gaps[:n_seg]
    return units, bits


def pack_gap_words(gap_nibbles):
    n = len(gap_nibbles)
    words = np.zeros(max((n + 7) // 8, 1), dtype=np.uint32)
    g = np.ascontiguousarray(gap_nibbles, dtype=np.uint8)
    lib().oracle_pack_gap_words(_ptr(g, _u8p), n, _ptr(words, _u32p))
    return words[:(n + 7) // 8]


def decode_flat(units, lut, n_symbols):
    units = np.ascontiguousarray(units, dtype=np.uint32)
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    lut_bits = int(lut.shape[0]).bit_length() - 1
    out = np.zeros(n_symbols, dtype=np.uint8)
    got = lib().oracle_decode_flat(_ptr(units, _u32p), len(units), lut.ctypes.data, lut_bits, _ptr(out, _u8p),
                                   n_symbols)
    return out[:got]


def sync_points(units, lut, n_symbols, subseq_bits=128):
    units = np.ascontiguousarray(units, dtype=np.uint32)
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    lut_bits = int(lut.shape[0]).bit_length() - 1
    n_sub = (len(units) * 32 + subseq_bits - 1) // subseq_bits
    off = np.zeros(max(n_sub, 1), dtype=np.uint8)
    cnt = np.zeros(max(n_sub, 1), dtype=np.uint32)
    lib().oracle_sync_points(_ptr(units, _u32p), len(units), lut.ctypes.data, lut_bits, n_symbols, subseq_bits,
                             _ptr(off, _u8p), _ptr(cnt, _u32p), n_sub)
    return off[:n_sub], cnt[:n_sub]


def decode_gap(units, gap_words, n_segments, lut, n_symbols):
    units = np.ascontiguousarray(units, dtype=np.uint32)
    gap_words = np.ascontiguousarray(gap_words, dtype=np.uint32)
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    lut_bits = int(lut.shape[0]).bit_length() - 1
    out = np.zeros(n_symbols, dtype=np.uint8)
    got = lib().oracle_decode_gap(_ptr(units, _u32p), len(units), _ptr(gap_words, _u32p), n_segments,
                                  lut.ctypes.data, lut_bits, _ptr(out, _u8p), n_symbols)
    return out[:got]


def huff_write(syms, lengths, original_size, units, gap_words, n_segments):
    syms = np.ascontiguousarray(syms, dtype=np.uint8)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
    units = np.ascontiguousarray(units, dtype=np.uint32)
    gap_words = np.ascontiguousarray(gap_words, dtype=np.uint32)
    size = lib().oracle_huff_size(len(syms), len(units), n_segments)
    buf = np.zeros(size, dtype=np.uint8)
    w = lib().oracle_huff_write(buf.ctypes.data, _ptr(syms, _u8p), _ptr(lengths, _u8p), len(syms), original_size,
                                _ptr(units, _u32p), len(units), _ptr(gap_words, _u32p), n_segments)
    assert w == size
    return buf


def huff_decode(file_bytes, out_cap):
    buf = np.frombuffer(bytes(file_bytes), dtype=np.uint8).copy()
    out = np.zeros(max(out_cap, 1), dtype=np.uint8)
    got = lib().oracle_huff_decode(buf.ctypes.data, len(buf), _ptr(out, _u8p), out_cap)
    if got < 0:
        raise ValueError("oracle_huff_decode failed: %d" % got)
    return out[:got]


# -------------------------------------------- simple length assignment (tests)
def huffman_lengths(data, max_len):
    """Length-limited code lengths for test inputs (plain Huffman, then a
    Kraft-sum repair if a length exceeds max_len).  Returns (syms, lengths)
    sorted by non-decreasing length -- the order both containers require."""
    import heapq
    hist = np.bincount(np.asarray(data, dtype=np.uint8), minlength=256)
    present = [int(s) for s in np.nonzero(hist)[0]]
    if len(present) == 1:
        return np.array(present, dtype=np.uint8), np.array([1], dtype=np.uint8)
    heap = [(int(hist[s]), i, [s]) for i, s in enumerate(present)]
    heapq.heapify(heap)
    depth = {s: 0 for s in present}
    uid = len(heap)
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        for s in a[2] + b[2]:
            depth[s] += 1
        heapq.heappush(heap, (a[0] + b[0], uid, a[2] + b[2]))
        uid += 1
    lens = {s: min(depth[s], max_len) for s in present}
    # repair Kraft inequality after clamping: lengthen the cheapest codes
    def kraft():
        return sum(1 << (max_len - l) for l in lens.values())
    order = sorted(present, key=lambda s: (hist[s], s))
    while kraft() > (1 << max_len):
        for s in order:
            if lens[s] < max_len:
                lens[s] += 1
                break
    # then use up slack by shortening the most frequent ones where possible
    for s in reversed(order):
        while lens[s] > 1 and kraft() + (1 << (max_len - lens[s])) <= (1 << max_len):
            lens[s] -= 1
    items = sorted(present, key=lambda s: (lens[s], s))
    return np.array(items, dtype=np.uint8), np.array([lens[s] for s in items], dtype=np.uint8)


# ----------------------------------------------------- reference-built checkers
def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_llhuff.so"))


_ref_llhuff = None
_ref_seq = None


def ref_llhuff_encode(data):
    """UNMODIFIED reference encoder (llhuffman_encoder.cc) -> dict with lut (2048x2 u8), units (incl. the
    appended zero unit of cuhd_input_buffer.cc:17-31), codes, lens."""
    global _ref_llhuff
    if _ref_llhuff is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_llhuff.so"))
        L.ref_llhuff_encode.restype = C.c_uint64
        L.ref_llhuff_encode.argtypes = [_u8p, C.c_uint64, _u8p, _u32p, C.c_uint64, _u32p, _u8p]
        _ref_llhuff = L
    data = np.ascontiguousarray(data, dtype=np.uint8)
    lut = np.zeros((2048, 2), dtype=np.uint8)
    cap = len(data) * 11 // 32 + 16
    units = np.zeros(cap, dtype=np.uint32)
    codes = np.zeros(256, dtype=np.uint32)
    lens = np.zeros(256, dtype=np.uint8)
    nu = _ref_llhuff.ref_llhuff_encode(_ptr(data, _u8p), len(data), _ptr(lut, _u8p), _ptr(units, _u32p), cap,
                                       _ptr(codes, _u32p), _ptr(lens, _u8p))
    if nu == 0:
        raise RuntimeError("reference encoder failed")
    return dict(lut=lut, units=units[:nu].copy(), codes=codes, lens=lens)


def ref_sequential_roundtrip(data):
    """UNMODIFIED reference CPU codec (sequential.cpp) -> (ok, decoded, compressed_bytes, enc_s, dec_s)."""
    global _ref_seq
    if _ref_seq is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_sequential.so"))
        L.ref_sequential_roundtrip.restype = C.c_int
        L.ref_sequential_roundtrip.argtypes = [_u8p, C.c_uint64, _u8p, C.POINTER(C.c_uint64),
                                               C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _ref_seq = L
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(len(data), dtype=np.uint8)
    cb = C.c_uint64(0)
    es = C.c_double(0)
    ds = C.c_double(0)
    ok = _ref_seq.ref_sequential_roundtrip(_ptr(data, _u8p), len(data), _ptr(out, _u8p), C.byref(cb), C.byref(es),
                                           C.byref(ds))
    return bool(ok), out, cb.value, es.value, ds.value
