"""Profiling target: a few decodes of one configuration (run it under ncu on the GPU box)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd

ap = argparse.ArgumentParser()
ap.add_argument("--scheme", default="gap")
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--r", type=float, default=0.5)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--segs", type=int, default=0)
ap.add_argument("--nt", type=int, default=0)
a = ap.parse_args()
torch.cuda.set_device(0)
hd.set_tuning(a.segs, a.nt)
n = a.mib << 20
data = hd.generate(n, a.r, seed=1)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
if a.scheme == "gap":
    pairs = hd.build_code(data, 16)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    tab = hd.Codetable(pairs=pairs)
    d_units = hd.to_device_units(units)
    d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
    dec = hd.Decoder(d_units.numel(), selfsync=False)
    fn = lambda: dec.decode_gap(d_units, d_gaps, tab, n, out=out)
else:
    pairs = hd.build_code(data, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    fn = lambda: dec.decode_selfsync(d_units, tab, n, out=out)
for _ in range(a.reps):
    fn()
torch.cuda.synchronize()
ok = bool(torch.equal(out, torch.from_numpy(data).cuda()))
print("ok", ok)
sys.exit(0 if ok else 1)
