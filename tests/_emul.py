"""ctypes door to tests/emul/libhd_emul.so: a CPU re-enactment of the CUDA kernels
that steps the product's own per-thread source (csrc/hd_core.h).  Test
infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cse375-finalproj-huffman-decoding_b200")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
_lib = None


def build():
    so = os.path.join(EMUL_DIR, "libhd_emul.so")
    srcs = [os.path.join(EMUL_DIR, "hd_emul.cpp"), os.path.join(PKG, "csrc", "hd_core.h"),
            os.path.join(PKG, "csrc", "hd_host.cpp"), os.path.join(PKG, "csrc", "hd_internal.h")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall",
                               "-I", os.path.join(PKG, "csrc"), "-I", os.path.join(ROOT, "include"),
                               srcs[0], srcs[2], "-o", so, "-lpthread"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emul_decode_gap.restype = C.c_int64
        L.emul_decode_gap.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                      C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32,
                                      C.c_uint32]
        L.emul_sync.restype = C.c_int64
        L.emul_sync.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                C.c_void_p, C.c_void_p, C.c_uint32]
        L.emul_table_info.restype = C.c_int
        L.emul_table_info.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _tbl(flat_lut, pairs):
    if flat_lut is not None:
        f = np.ascontiguousarray(flat_lut, dtype=np.uint8)
        return f, f.ctypes.data, int(f.shape[0]).bit_length() - 1, None, None, 0
    p = np.ascontiguousarray(pairs, dtype=np.uint8)
    return None, None, 0, p, p.ctypes.data, p.shape[0]


def decode_gap(units, gap_words, n_segments, n_symbols, flat_lut=None, pairs=None, first_bit=0, phase=0, segs=None,
               nt=32):
    if segs is None:   # exercise every tile geometry the kernel is instantiated for
        res = [decode_gap(units, gap_words, n_segments, n_symbols, flat_lut, pairs, first_bit, phase, s, n)
               for s, n in ((1, 32), (2, 32), (4, 32))]
        res = [r for r in res if r is not None]   # (2, 32) / (4, 32) do not exist for short minimum code lengths
        assert all(np.array_equal(res[0], r) for r in res[1:]), "chunk geometries disagree"
        return res[0]
    units = np.ascontiguousarray(units, dtype=np.uint32)
    gw = np.ascontiguousarray(gap_words, dtype=np.uint32)
    f, fp, fb, p, pp, npairs = _tbl(flat_lut, pairs)
    out = np.zeros(n_symbols + 64, dtype=np.uint8)
    got = lib().emul_decode_gap(units.ctypes.data, len(units), gw.ctypes.data, n_segments, first_bit, fp, fb, pp,
                                npairs, out.ctypes.data, n_symbols, phase, segs, nt)
    if got == -40 and nt == 32 and segs != 1:
        return None
    if got < 0:
        raise RuntimeError("emul_decode_gap failed: %d" % got)
    return out[:got]


def sync(units, flat_lut=None, pairs=None, first_bit=0, exact=False):
    """-> (gap words, K1b rounds run, [tiles whose published exit moved in K1a, K1b rounds])."""
    units = np.ascontiguousarray(units, dtype=np.uint32)
    f, fp, fb, p, pp, npairs = _tbl(flat_lut, pairs)
    n_seg = (len(units) + 3) // 4
    n_chunks = (n_seg + 15) // 16
    gw = np.zeros(max(2 * n_chunks, 1), dtype=np.uint32)
    rounds = np.zeros(4, dtype=np.uint32)
    r = lib().emul_sync(units.ctypes.data, len(units), first_bit, fp, fb, pp, npairs, gw.ctypes.data,
                        rounds.ctypes.data, 1 if exact else 0)
    if r < 0:
        raise RuntimeError("emul_sync failed: %d" % r)
    return gw, int(r), rounds[:2].tolist()


def table_info(flat_lut=None, pairs=None):
    f, fp, fb, p, pp, npairs = _tbl(flat_lut, pairs)
    info = np.zeros(4, dtype=np.uint32)
    if lib().emul_table_info(fp, fb, pp, npairs, info.ctypes.data):
        raise RuntimeError("emul_table_info failed")
    return dict(lut_bits=int(info[0]), min_len=int(info[1]), max_len=int(info[2]), has_long=int(info[3]))
