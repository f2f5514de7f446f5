"""End-to-end (pinned host in/out) timing of hd_host_decode_selfsync with its internal timing triple."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd

torch.cuda.set_device(0)
n = 100 << 20
data = hd.generate(n, 0.5, seed=1)
pairs = hd.build_code(data, 11)
units, bits = hd.encode(data, pairs, pad_unit=True)
lut = hd.flat_lut(pairs, 11)
h_units = torch.from_numpy(units.view(np.int32)).pin_memory()
h_out = torch.zeros(n, dtype=torch.uint8).pin_memory()
dec = hd.HostDecoder(len(units), n)
for _ in range(3):
    dec.decode_selfsync(h_units, lut, h_out, n)
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    tm = dec.decode_selfsync(h_units, lut, h_out, n)
    ts.append((time.perf_counter() - t0, tm))
print("  all walls ms:", " ".join("%.2f" % (x[0] * 1e3) for x in ts))
ts.sort(key=lambda x: x[0])
t, tm = ts[len(ts) // 2]
print("piece log2 %s: wall %.3f ms  %.1f GB/s  [first piece up %.3f, decode span %.3f, download tail %.3f]  ok=%s" % (
    os.environ.get("HD_HOST_PIECE_LOG2", "20"), t * 1e3, n / t / 1e9, tm[0], tm[1], tm[2], bool(np.array_equal(h_out.numpy(), data))))
