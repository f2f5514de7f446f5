"""Randomised GPU round trips (sizes, redundancies, code-length limits, both schemes, host pipeline): developer check."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd

torch.cuda.set_device(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 40.0)
n_cases = 0
while time.time() < t_end:
    n = int(rng.choice([rng.integers(1, 5000), rng.integers(5000, 400000), rng.integers(400000, 6000000)]))
    r = float(rng.choice([0.0, 0.1, 0.3, 0.5, 0.7, 0.9, 0.97, 1.0]))
    data = hd.generate(n, r, seed=int(rng.integers(1, 1 << 30)))
    if rng.random() < 0.2:      # runs of one symbol: worst case for the staging rows
        k = int(rng.integers(1, max(2, n // 2)))
        s = int(rng.integers(0, n - k + 1))
        data[s:s + k] = data[s]
    ref = torch.from_numpy(data).cuda()
    # gap scheme, code lengths up to 16
    lim = int(rng.integers(9, 17))
    pairs = hd.build_code(data, lim)
    if rng.random() < 0.3:      # an INCOMPLETE code (Kraft sum < 1: LUT holes, resolver variants of the kernels): lengthen
        lens = pairs[:, 1].astype(np.int64)           # some codewords, keep the list order non-decreasing
        lens = np.minimum(16, lens + rng.integers(0, 3, len(lens)))
        pairs = pairs.copy()
        pairs[:, 1] = np.maximum.accumulate(lens).astype(np.uint8)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    tab = hd.Codetable(pairs=pairs)
    d_units = hd.to_device_units(units)
    d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
    dec = hd.Decoder(max(d_units.numel(), 1), selfsync=False)
    out = dec.decode_gap(d_units, d_gaps, tab, n)
    assert torch.equal(out, ref), ("gap", n, r, lim)
    tab.close()
    # self-sync scheme, flat 11-bit LUT
    pairs = hd.build_code(data, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    lut = hd.flat_lut(pairs, 11)
    tab = hd.Codetable(flat_lut=lut)
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(max(d_units.numel(), 1), selfsync=True)
    out = dec.decode_selfsync(d_units, tab, n)
    assert torch.equal(out, ref), ("selfsync", n, r)
    tab.close()
    if n > 100000 and rng.random() < 0.5:
        hdec = hd.HostDecoder(len(units), n)
        h_out = np.zeros(n, dtype=np.uint8)
        hdec.decode_selfsync(units, lut, h_out, n)
        assert np.array_equal(h_out, data), ("host", n, r)
        hdec.close()
    n_cases += 1
print("fuzz ok:", n_cases, "cases")
