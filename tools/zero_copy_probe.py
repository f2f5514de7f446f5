"""Does K4 sustain PCIe write bandwidth when its output pointer is pinned HOST memory (UVA)?  Developer probe."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd

torch.cuda.set_device(0)
n = 100 << 20
data = hd.generate(n, 0.5, seed=1)
pairs = hd.build_code(data, 16)
units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
tab = hd.Codetable(pairs=pairs)
d_units = hd.to_device_units(units)
d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
dec = hd.Decoder(d_units.numel(), selfsync=False)
h_out = torch.zeros(n, dtype=torch.uint8).pin_memory()
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, out in (("device out", d_out), ("pinned host out", h_out)):
    for _ in range(2):
        dec.decode_gap(d_units, d_gaps, tab, n, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dec.decode_gap(d_units, d_gaps, tab, n, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[2]
    ok = bool(np.array_equal(out.cpu().numpy(), data))
    print("%-16s %.3f ms  %.1f GB/s  ok=%s" % (name, t, n / t / 1e6, ok))
# with a concurrent H2D of the same stream size on another stream
s2 = torch.cuda.Stream()
h_in = torch.empty(78 << 20, dtype=torch.uint8).pin_memory(); d_in = torch.empty(78 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s2):
    d_in.copy_(h_in, non_blocking=True)
dec.decode_gap(d_units, d_gaps, tab, n, out=h_out)
torch.cuda.synchronize(); t = time.perf_counter() - t0
print("pinned host out + concurrent 78 MiB H2D: %.3f ms  (%.1f GB/s decoded)" % (t * 1e3, n / t / 1e9))
