"""Small end-to-end run for compute-sanitizer (memcheck): both schemes, odd sizes, unaligned output, 16-bit codes,
device-built tables and the device encoder."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hd_b200 as hd
import cases

torch.cuda.set_device(0)
ok = True
for n, r in ((1, 0.5), (4097, 0.5), (300_001, 0.5), (200_003, 1.0), (150_000, 0.0)):
    data = hd.generate(n, r, seed=3)
    pairs = hd.build_code(data, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    buf = torch.zeros(n + 33, dtype=torch.uint8, device="cuda")
    out = dec.decode_selfsync(d_units, tab, n, out=buf[3:3 + n])
    ok &= bool(np.array_equal(out.cpu().numpy(), data))
    gw = dec.gap_array(d_units, tab)
    pairs16 = hd.build_code(data, 16)
    units16, gaps, _ = hd.encode(data, pairs16, want_gaps=True)
    tab16 = hd.Codetable(device_pairs=torch.from_numpy(np.ascontiguousarray(pairs16)).cuda())
    d_u = hd.to_device_units(units16)
    d_g = torch.from_numpy(gaps.view(np.int32)).cuda()
    dec2 = hd.Decoder(d_u.numel(), selfsync=False)
    ok &= bool(np.array_equal(dec2.decode_gap(d_u, d_g, tab16, n).cpu().numpy(), data))
    p2, du2, dg2, b2 = hd.encode_device(torch.from_numpy(data).cuda(), 16)
    ok &= bool(np.array_equal(du2.cpu().numpy().view(np.uint32), units16))
    tab.close(); tab16.close()
pairs = cases.skewed_long_code()
data = cases.data_for_pairs(pairs, 100_003, seed=9, geometric=True)
units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
tab = hd.Codetable(pairs=pairs)
d_units = hd.to_device_units(units)
dec = hd.Decoder(d_units.numel(), selfsync=True)
ok &= bool(np.array_equal(dec.decode_selfsync(d_units, tab, len(data)).cpu().numpy(), data))
pairs = cases.fixed_length_code(8, 3)
data = cases.data_for_pairs(pairs, 150_001, seed=5)
units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
tab = hd.Codetable(flat_lut=lut)
d_units = hd.to_device_units(units)
dec = hd.Decoder(d_units.numel(), selfsync=True)
ok &= bool(np.array_equal(dec.decode_selfsync(d_units, tab, len(data)).cpu().numpy(), data))
print("ok", ok)
sys.exit(0 if ok else 1)
