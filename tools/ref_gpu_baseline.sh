#!/bin/bash
# Same-box GPU baselines: the UNMODIFIED reference GPU programs rebuilt for sm_100 (oracle/_ref, oracle/Makefile) on
# 100 MiB of generate.cpp-distribution data at redundancy 0.5 and 0.9.
set -u
R=$(pwd)
W=$(mktemp -d)
mkdir -p $W/run gpurun_out
for r in 0.5 0.9; do
  python - <<PY
import sys; sys.path.insert(0, "$R")
import hd_b200 as hd
hd.generate(100 << 20, $r, seed=1).tofile("$W/data100_100.bin")
PY
  echo "=== redundancy $r: reference gpuhd demo (self-sync; 'Decompression time' includes malloc + H2D + D2H) ==="
  (cd $W/run && timeout 300 $R/oracle/_ref/gpuhd_demo 0 100) 2>&1 | tail -8
  echo "=== redundancy $r: our bin/demo on the same file ==="
  (cd $W/run && timeout 300 $R/cse375-finalproj-huffman-decoding_b200/bin/demo 0 100) 2>&1 | tail -6
  echo "=== redundancy $r: reference gap-array encoder + decoder ('dec,' = kernel ms, mean of 100 runs) ==="
  (cd $W && timeout 300 $R/oracle/_ref/gap_encoder data100_100.bin ref.huff 2>&1 | tail -4; timeout 600 $R/oracle/_ref/gap_decoder ref.huff decoded_ref 2>&1 | tail -5; cmp decodedfile data100_100.bin && echo "reference decoder output: identical" || echo "reference decoder output: DIFFERS")
  echo "=== redundancy $r: our bin/decoder on the reference's .huff ==="
  (cd $W && timeout 300 $R/cse375-finalproj-huffman-decoding_b200/bin/decoder ref.huff decoded_ours 2>&1 | tail -4; cmp decoded_ours data100_100.bin && echo "our decoder output: identical")
done
rm -rf $W
