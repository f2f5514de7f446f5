#!/bin/bash
mkdir -p gpurun_out
R=$(pwd)
python tools/quick_bench.py --mib 100 --r 0.9 0.95 --iters 5 2>&1 | tee gpurun_out/qb_r09.log
W=$(mktemp -d); mkdir -p $W/run
python - <<PY
import sys; sys.path.insert(0, "$R")
import hd_b200 as hd, numpy as np
d = hd.generate(100 << 20, 0.9, seed=1); d.tofile("$W/data100_100.bin")
print("hist top", np.sort(np.bincount(d, minlength=256))[::-1][:8])
PY
(cd $W/run && $R/cse375-finalproj-huffman-decoding_b200/bin/demo 0 100; $R/cse375-finalproj-huffman-decoding_b200/bin/demo 0 100) 2>&1 | grep -i "time\|mismatch"
(cd $W && $R/oracle/_ref/gap_encoder data100_100.bin ref.huff > /dev/null 2>&1)
python - <<PY
import sys, time; sys.path.insert(0, "$R")
import hd_b200 as hd, numpy as np, torch
img = np.fromfile("$W/ref.huff", dtype=np.uint8)
n = int(np.frombuffer(img[:8].tobytes(), dtype=np.uint64)[0])
pl = np.frombuffer(img[8:8 + 2 * n].tobytes(), dtype=np.uint8).reshape(n, 2)
print("reference .huff: symbols", n, "max len", pl[:, 1].max(), "min len", pl[:, 1].min(), "len hist", np.bincount(pl[:, 1]))
for i in range(3):
    t = time.time(); out = hd.decode_huff(img); print("decode_huff s", time.time() - t)
PY
rm -rf $W
