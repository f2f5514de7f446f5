"""> 4 GiB round trip on the device (the reference's formats stop at 4 GiB): device encoder -> gap decode and
self-sync decode -> compare.  Developer probe; the pytest version is tests/test_gpu_parity.py::test_beyond_4_gib."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd

torch.cuda.set_device(0)
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.5
n = int(gib * (1 << 30)) + 12345
g = torch.Generator(device="cuda"); g.manual_seed(7)
# skewed bytes: half of the mass on 4 symbols (the generate.cpp shape), the rest uniform
t0 = time.time()
data = torch.empty(n, dtype=torch.uint8, device="cuda")
step = 1 << 28
for o in range(0, n, step):
    m = min(step, n - o)
    u = torch.randint(0, 256, (m,), dtype=torch.int16, device="cuda", generator=g)
    hot = torch.randint(0, 2, (m,), dtype=torch.int16, device="cuda", generator=g)
    data[o:o + m] = torch.where(hot == 1, u & 3, u).to(torch.uint8)
    del u, hot
torch.cuda.synchronize(); print("generated %.2f GiB in %.1fs" % (n / 2**30, time.time() - t0), flush=True)
for max_len, scheme in ((16, "gap"), (11, "selfsync")):
    t0 = time.time()
    pairs, d_units, d_gaps, bits = hd.encode_device(data, max_len, pad_unit=(scheme == "selfsync"))
    torch.cuda.synchronize(); print(scheme, "encoded: %.2f GiB of stream, %d bits, %.1fs" % (d_units.numel() * 4 / 2**30, bits, time.time() - t0), flush=True)
    tab = hd.Codetable(pairs=pairs) if scheme == "gap" else hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
    dec = hd.Decoder(d_units.numel(), selfsync=(scheme == "selfsync"))
    out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        a.record()
        if scheme == "gap":
            dec.decode_gap(d_units, d_gaps, tab, n, out=out)
        else:
            dec.decode_selfsync(d_units, tab, n, out=out)
        b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    ok = bool(torch.equal(out, data))
    print("%s decode of %.2f GiB: %.2f ms (%.0f GB/s) identical=%s" % (scheme, n / 2**30, ms, n / ms / 1e6, ok), flush=True)
    del out, dec, tab, d_units, d_gaps
