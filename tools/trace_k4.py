"""Developer aid: per-step timeline of one CTA of the decode-and-write kernel (hd_set_trace)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd
ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=100)
ap.add_argument("--segs", type=int, default=2)
ap.add_argument("--nt", type=int, default=512)
a = ap.parse_args()
torch.cuda.set_device(0)
hd.set_tuning(a.segs, a.nt)
n = a.mib << 20
data = hd.generate(n, 0.5, seed=1)
pairs = hd.build_code(data, 16)
units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
tab = hd.Codetable(pairs=pairs)
d_units = hd.to_device_units(units)
d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
dec = hd.Decoder(d_units.numel(), selfsync=False)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dec.decode_gap(d_units, d_gaps, tab, n, out=out)
trace = torch.zeros(48 * 8, dtype=torch.int64, device="cuda")
hd.lib().hd_set_trace(trace.data_ptr())
dec.decode_gap(d_units, d_gaps, tab, n, out=out)
torch.cuda.synchronize()
hd.lib().hd_set_trace(None)
tr = trace.cpu().numpy().reshape(48, 8)
t0 = tr[0, 0]
print("step  dec_start loaded   dec_end  lb_polls     step_end | full_passed lookback_done tile   (us from start)")
for k in range(48):
    if tr[k, 0] == 0:
        break
    f = lambda v: "%8.2f" % ((v - t0) / 1e3) if v else "       -"
    print("%3d  %s %s %s %s %s | %s %s %6d" % (k, f(tr[k, 0]), f(tr[k, 7]), f(tr[k, 1]), "%8d" % tr[k, 2], f(tr[k, 3]), f(tr[k, 4]), f(tr[k, 5]), tr[k, 6]))
