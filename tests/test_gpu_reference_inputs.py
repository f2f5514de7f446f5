"""Parity AT SCALE on inputs made by the reference's OWN producers: the generator
(generate.cpp -> oracle/_ref/generate), the gpuhd llhuff encoder (oracle/_ref/libref_llhuff.so) and the gap-array GPU
encoder (oracle/_ref/gap_encoder), all built unmodified by oracle/Makefile and shipped to the GPU box.  The CUDA path
runs through the C ABI; results are compared with the original bytes (D1-aware where the reference encoder drops the
tail of the last codeword, llhuffman_encoder.cc:213-237) and, at sizes the oracle finishes in seconds, with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as O
import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "cse375-finalproj-huffman-decoding_b200", "bin")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


def _ref_generate(tmp_path, n, r):
    """n bytes from the unmodified generate.cpp binary (writes ./data.bin, unseeded: generate.cpp:24,32)."""
    exe = os.path.join(REF, "generate")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/generate not built (needs /root/reference at build time)")
    procs, piece = 8, (n + 7) // 8
    jobs = []
    for i in range(procs):
        size = min(piece, n - i * piece)
        d = tmp_path / ("g%d" % i)
        d.mkdir()
        jobs.append((d, size, subprocess.Popen([exe, str(size), repr(float(r))], cwd=d, stdout=subprocess.DEVNULL)))
    parts = []
    for d, size, p in jobs:
        assert p.wait() == 0
        parts.append(np.fromfile(d / "data.bin", dtype=np.uint8, count=size))
        os.remove(d / "data.bin")
    return np.concatenate(parts)


def _d1_equal(got, data):
    return np.array_equal(got, data) or (len(got) == len(data) and np.array_equal(got[:-1], data[:-1]))


def test_100MiB_stream_from_the_reference_llhuff_encoder(hd, torch_cuda, tmp_path):
    """configs[1] with BOTH producers from the reference: generate.cpp data, llhuff tables + encode_memory + pad unit;
    decode == original (D1-aware) and == the oracle's decode of the same stream (exactly, incl. the last symbol)."""
    if not O.have_ref():
        pytest.skip("oracle/_ref/libref_llhuff.so not built")
    n = 100 << 20
    data = _ref_generate(tmp_path, n, 0.5)
    hist = np.bincount(data, minlength=256)
    assert hist[65:69].sum() > 0.45 * n                           # the generator's hot symbols 'A'..'D'
    ref = O.ref_llhuff_encode(data)
    got = hd.decode_selfsync(ref["units"], ref["lut"], n)
    assert _d1_equal(got, data)
    want = O.decode_flat(ref["units"], ref["lut"], n)             # CPU oracle on the reference's stream
    assert np.array_equal(got, want)
    # the host-buffer pipeline (what bin/demo uses) on the same reference stream
    torch = torch_cuda
    hdec = hd.HostDecoder(len(ref["units"]), n, device=0)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    hdec.decode_selfsync(torch.from_numpy(ref["units"].view(np.int32)).pin_memory(), ref["lut"], h_out, n)
    assert np.array_equal(h_out.numpy(), want)
    hdec.close()


@pytest.mark.parametrize("r", [0.5, 0.9])
def test_1GiB_huff_from_the_reference_gpu_encoder(hd, torch_cuda, tmp_path, r):
    """configs[2]: a 1 GiB .huff written by the UNMODIFIED reference GPU encoder, decoded through the library (device
    path and pipelined host path) and through bin/decoder, equals the original file."""
    enc = os.path.join(REF, "gap_encoder")
    if not os.path.exists(enc):
        pytest.skip("oracle/_ref/gap_encoder not built")
    torch = torch_cuda
    n = 1 << 30
    data = _ref_generate(tmp_path, n, r)
    src, dst = tmp_path / "in.bin", tmp_path / "ref.huff"
    data.tofile(src)
    run = subprocess.run([enc, str(src), str(dst)], capture_output=True, text=True)
    if run.returncode != 0 or not dst.exists():
        pytest.skip("the reference encoder does not run on this GPU: " + (run.stderr or run.stdout)[-200:])
    os.remove(src)
    image = np.fromfile(dst, dtype=np.uint8)
    hf = hd.HuffFile(image)
    assert hf.original_size == n and hf.n_segments == (hf.n_units + 3) // 4
    tab = hd.Codetable(pairs=hf.pairs)
    d_units = hd.to_device_units(hf.units)
    d_gaps = torch.from_numpy(hf.gap_words.view(np.int32)).cuda()
    dec = hd.Decoder(d_units.numel(), selfsync=False)
    out = dec.decode_gap(d_units, d_gaps, tab, n)
    ref_t = torch.from_numpy(data).cuda()
    assert bool(torch.equal(out, ref_t)) and int(dec.count.item()) == n
    tab.close()
    del out, dec, d_units, d_gaps
    hdec = hd.HostDecoder(hf.n_units, n, device=0)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    size, _ = hdec.decode_huff(torch.from_numpy(image).pin_memory(), h_out)
    assert size == n and bool(torch.equal(h_out.cuda(), ref_t))
    hdec.close()
    del h_out, ref_t
    exe = os.path.join(BIN, "decoder")
    if os.path.exists(exe):
        outp = tmp_path / "out.bin"
        subprocess.run([exe, str(dst), str(outp)], check=True, capture_output=True)
        back = np.fromfile(outp, dtype=np.uint8)
        assert np.array_equal(back, data)


@pytest.mark.parametrize("r", cases.SWEEP)
def test_1GiB_sweep_both_schemes(hd, torch_cuda, r):
    """configs[4]: the whole redundancy sweep at 1 GiB, self-sync (11-bit code, + the self-sync -> gap-array transcode
    decoded through the gap path) and gap-array (16-bit code); byte-exact against the original and counts == n."""
    torch = torch_cuda
    n = 1 << 30
    data = hd.generate(n, r, seed=int(1000 * r) + 3)
    ref_t = torch.from_numpy(data).cuda()
    pairs = hd.build_code(data, 11)
    units, _ = hd.encode(data, pairs, pad_unit=True)
    tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    out = dec.decode_selfsync(d_units, tab, n)
    assert bool(torch.equal(out, ref_t)) and int(dec.count.item()) == n
    out.zero_()
    gaps = dec.gap_array(d_units, tab)                       # EXACT landings = the .huff gap array of this stream
    dec.decode_gap(d_units, gaps, tab, n, out=out)
    assert bool(torch.equal(out, ref_t))
    tab.close()
    del dec, d_units, gaps, units
    pairs16 = hd.build_code(data, 16)
    units16, gw, _ = hd.encode(data, pairs16, want_gaps=True)
    tab16 = hd.Codetable(pairs=pairs16)
    d_units = hd.to_device_units(units16)
    d_gaps = torch.from_numpy(gw.view(np.int32)).cuda()
    dec = hd.Decoder(d_units.numel(), selfsync=False)
    out.zero_()
    dec.decode_gap(d_units, d_gaps, tab16, n, out=out)
    assert bool(torch.equal(out, ref_t)) and int(dec.count.item()) == n
    tab16.close()


@pytest.mark.parametrize("r,n", [(0.5, 1_000_003), (0.9, 4_000_000), (0.0, 70_001), (1.0, 333_333)])
def test_device_encoder_payload_equals_the_oracle_packer(hd, torch_cuda, r, n):
    """f2 chain closed: the device encoder's payload and gap array against the ORACLE packer directly
    (oracle/hd_oracle.c restating encoder.cu:281-347, 358-379), not via the host product encoder."""
    torch = torch_cuda
    data = O.generate(n, r, seed=int(r * 100) + 5)
    pairs, d_units, d_gaps, bits = hd.encode_device(torch.from_numpy(data).cuda(), 16)
    codes, lens = O.canonical_codes(pairs[:, 0], pairs[:, 1])
    o_units, o_bits, o_nibbles = O.pack(data, codes, lens, flush_tail=True, want_gaps=True)
    assert bits == o_bits
    got_units = d_units.cpu().numpy().view(np.uint32)
    assert np.array_equal(got_units[:len(o_units)], o_units) and not got_units[len(o_units):].any()
    got_gaps = d_gaps.cpu().numpy().view(np.uint32)
    want_gaps = O.pack_gap_words(o_nibbles)
    assert np.array_equal(got_gaps[:len(want_gaps)], want_gaps)
