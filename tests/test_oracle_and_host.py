"""CPU-only: (1) the oracle is pinned to the reference -- committed golden vectors produced by the UNMODIFIED reference
encoder, plus the live reference build when oracle/_ref exists; (2) the host side of the library (encoder, containers,
sharding arithmetic) agrees with the oracle bit for bit; (3) the C-ABI library loads and exports every symbol that
include/hd_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import _oracle as O
import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ (1) oracle pinned to the reference
def test_oracle_decodes_reference_encoder_golden_vectors():
    gold, meta = cases.golden()
    assert len(gold) >= 10
    for name, c in gold.items():
        m = c["meta"]
        assert hashlib.sha256(c["data"].tobytes()).hexdigest() == m["sha256_data"], name
        assert hashlib.sha256(c["units"].tobytes()).hexdigest() == m["sha256_units"], name
        assert m["reference_sequential_pass"], name            # sequential.cpp round trip PASSed on these bytes
        dec = O.decode_flat(c["units"], c["lut"], len(c["data"]))
        if m["oracle_equals_original"]:
            assert np.array_equal(dec, c["data"]), name
        else:                                                   # reference packer defect D1: last codeword truncated
            assert np.array_equal(dec[:-1], c["data"][:-1]), name


def test_generator_restatement_is_reproducible():
    _, meta = cases.golden()
    for key, sha in meta["generator_1MiB"].items():
        r, seed = int(key[1:4]) / 100.0, int(key.split("_s")[1])
        assert hashlib.sha256(O.generate(1 << 20, r, seed).tobytes()).hexdigest() == sha, key


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_oracle_packer_and_lut_equal_live_reference_encoder():
    for n, r, seed in ((1000, 0.5, 4), (4099, 0.5, 5), (3000, 0.0, 6), (3001, 0.9, 7), (2000, 1.0, 8)):
        d = O.generate(n, r, seed)
        ref = O.ref_llhuff_encode(d)
        codes, lens = ref["codes"], ref["lens"]
        assert np.array_equal(O.flat_lut(codes, lens, 11), ref["lut"])
        units, bits, _ = O.pack(d, codes, lens, flush_tail=True, want_gaps=True)
        ru = ref["units"]
        assert len(ru) == len(units) + 1 and ru[-1] == 0                               # the appended zero unit
        # everything before the final data unit is identical; the final unit itself depends on the reference packer's
        # end-of-input handling (defects D1/D2, llhuffman_encoder.cc:213-237: a straddling last codeword loses its
        # tail, and the flush relies on an underflowing shift count), so it is only checked through the decode
        assert np.array_equal(units[:-1], ru[:-2])
        dec = O.decode_flat(ru, ref["lut"], n)
        assert np.array_equal(dec[:-1], d[:-1])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_reference_sequential_roundtrip_on_oracle_generator_data():
    d = O.generate(50_000, 0.5, 3)
    ok, dec, *_ = O.ref_sequential_roundtrip(d)
    assert ok and np.array_equal(dec, d)


# ------------------------------------------------------------------ (2) host side vs oracle
@pytest.mark.parametrize("r", [0.0, 0.5, 0.9, 1.0])
@pytest.mark.parametrize("limit", [11, 16])
def test_host_encoder_matches_oracle_packer(hd, r, limit):
    n = 70_001
    data = O.generate(n, r, seed=int(r * 10) + limit)
    pairs = hd.build_code(data, limit)
    assert (np.diff(pairs[:, 1].astype(int)) >= 0).all() and pairs[:, 1].max() <= limit
    assert sum(2.0 ** -int(l) for l in pairs[:, 1]) <= 1.0 + 1e-12                    # Kraft
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    codes, lens = O.canonical_codes(pairs[:, 0], pairs[:, 1])
    units_o, bits_o, gap_o = O.pack(data, codes, lens, flush_tail=True, want_gaps=True)
    assert bits == bits_o and np.array_equal(units, units_o)
    assert np.array_equal(gaps, O.pack_gap_words(gap_o))
    lut_bits = int(pairs[:, 1].max())
    assert np.array_equal(hd.flat_lut(pairs, lut_bits), O.flat_lut(codes, lens, lut_bits))
    # containers: byte-identical .huff image, and the parser gives back what went in
    n_seg = (len(units) + 3) // 4
    image = hd.huff_write(pairs, n, units, gaps)
    image_o = O.huff_write(pairs[:, 0], pairs[:, 1], n, units, gaps, n_seg)
    assert np.array_equal(image, image_o)
    hf = hd.HuffFile(image)
    assert hf.original_size == n and hf.n_units == len(units) and hf.n_segments == n_seg
    assert np.array_equal(hf.units, units) and np.array_equal(hf.pairs, pairs)
    assert np.array_equal(O.huff_decode(image, n), data)


def test_huff_parser_rejects_malformed_images(hd):
    data = O.generate(5000, 0.5, seed=1)
    pairs = hd.build_code(data, 16)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    image = hd.huff_write(pairs, len(data), units, gaps)
    for bad in (image[:7], image[:len(image) // 2], image[:-1], np.zeros(4, np.uint8)):
        with pytest.raises(hd.HDError):
            hd.HuffFile(bad)


# ------------------------------------------------------------------ (3) the C ABI
def test_library_exports_every_symbol_the_header_declares(hd):
    hdr = open(os.path.join(ROOT, "include", "hd_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hd_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    lib = C.CDLL(hd.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert sorted(hd.ABI_SYMBOLS) == declared                      # the Python mirror binds the same list
    assert b"sm_100a" in lib_version(lib)


def lib_version(lib):
    lib.hd_version.restype = C.c_char_p
    return lib.hd_version()


def test_status_strings_and_pure_host_entries(hd):
    L = hd.lib()
    assert L.hd_status_string(0) == b"ok" and L.hd_status_string(4) == b"malformed input"
    assert L.hd_workspace_bytes_selfsync(1 << 20) > L.hd_workspace_bytes_gap(1 << 20) > 0
    assert L.hd_workspace_bytes_gap(1 << 24) > L.hd_workspace_bytes_gap(1 << 20)
    # workspace is a few percent of the stream (the reference needs 20 B per 16 B subsequence)
    assert L.hd_workspace_bytes_selfsync(1 << 24) < 0.05 * 4 * (1 << 24)


def test_no_cpu_fallback(hd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    data = O.generate(1000, 0.5, seed=2)
    pairs = hd.build_code(data, 11)
    with pytest.raises(hd.HDError):
        hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
    with pytest.raises(hd.HDError):
        hd.decode_selfsync(np.zeros(8, np.uint32), hd.flat_lut(pairs, 11), 10)
    # and straight through the C ABI: table upload needs a device
    h = C.c_void_p(None)
    lut = np.ascontiguousarray(hd.flat_lut(pairs, 11))
    st = hd.lib().hd_table_from_flat_lut(lut.ctypes.data, 11, None, C.byref(h))
    assert st == 2 and h.value is None                              # HD_ERR_CUDA
