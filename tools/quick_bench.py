"""Developer timing loop (not the contract bench): device-timed decode throughput for both schemes."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd


def timed(fn, iters, flush, fin=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        flush.add_(1)          # > L2 write between iterations
        ev[i][0].record()
        fn()
        ev[i][1].record()
        if fin:
            fin()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, nargs="+", default=[100, 1024])
    ap.add_argument("--r", type=float, nargs="+", default=[0.5])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--schemes", nargs="+", default=["gap", "selfsync"])
    ap.add_argument("--geom", nargs="+", default=["0,0"], help="segs,nt pairs (0,0 = automatic)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    for mib in a.mib:
        for r in a.r:
            n = mib << 20
            t0 = time.time()
            data = hd.generate(n, r, seed=1)
            t1 = time.time()
            ref = torch.from_numpy(data).cuda()
            for scheme in a.schemes:
                if scheme == "gap":
                    pairs = hd.build_code(data, 16)
                    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
                    tab = hd.Codetable(pairs=pairs)
                    d_units = hd.to_device_units(units)
                    d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
                    dec = hd.Decoder(d_units.numel(), selfsync=False)
                    out = torch.empty(n, dtype=torch.uint8, device="cuda")
                    fn = lambda: dec.decode_gap(d_units, d_gaps, tab, n, out=out)
                    extra = 4 * len(gaps)
                else:
                    pairs = hd.build_code(data, 11)
                    units, bits = hd.encode(data, pairs, pad_unit=True)
                    tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
                    d_units = hd.to_device_units(units)
                    dec = hd.Decoder(d_units.numel(), selfsync=True)
                    out = torch.empty(n, dtype=torch.uint8, device="cuda")
                    fn = lambda: dec.decode_selfsync(d_units, tab, n, out=out, defer=True)
                    extra = 0
                for geom in a.geom:
                    segs, nt = map(int, geom.split(","))
                    hd.set_tuning(segs, nt)
                    out.zero_()
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    ok = bool(torch.equal(out, ref))
                    fin = dec.finish if scheme == "selfsync" else None
                    if fin:
                        fin()
                    med, best = timed(fn, a.iters, flush, fin)
                    hd.profile_enable(True)
                    fn()
                    if fin:
                        fin()
                    prof = hd.profile_read()
                    hd.profile_enable(False)
                    algo = 4 * len(units) + n + extra
                    print("%-8s %5d MiB r=%.2f geom=%s ok=%s  median %.3f ms  best %.3f ms  decoded %.1f GB/s  "
                          "roofline-bytes %.1f GB/s  kernels %s" % (scheme, mib, r, geom, ok, med, best, n / med / 1e6,
                                                                   algo / med / 1e6,
                                                                   {k: round(v, 3) for k, v in prof.items() if v}),
                          flush=True)
                tab.close()


if __name__ == "__main__":
    main()
