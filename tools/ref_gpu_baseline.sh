#!/bin/bash
# Same-box GPU baselines: the UNMODIFIED reference GPU programs rebuilt for sm_100 (oracle/_ref, oracle/Makefile) next to
# our CLIs on the same 100 MiB files from the reference's generate.cpp at redundancy 0.5 and 0.9 (BASELINE.md section 3,
# step 4).  Our CLIs run twice: the first process pays the CUDA context + module load, like the reference's.
set -u
R=$(pwd)
B=$R/cse375-finalproj-huffman-decoding_b200/bin
W=$(mktemp -d)
mkdir -p $W/run gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for r in 0.5 0.9; do
  (cd $W && $R/oracle/_ref/generate 104857600 $r > /dev/null && mv data.bin data100_100.bin && cp data100_100.bin input_data.bin)
  echo "=== redundancy $r: reference gpuhd demo (self-sync; 'Decompression time' = malloc + H2D + decode + D2H) ==="
  (cd $W/run && timeout 300 $R/oracle/_ref/gpuhd_demo 0 100) 2>&1 | tail -7
  echo "--- four more processes of the reference demo (the timed region holds four cudaMalloc calls: single runs vary by tens of ms) ---"
  (cd $W/run && for i in 1 2 3 4; do timeout 300 $R/oracle/_ref/gpuhd_demo 0 100 2>&1 | grep -E "Decompression time|mismatch" | tr '\n' ' '; echo; done)
  echo "=== redundancy $r: our bin/demo on the same file (five processes) ==="
  (cd $W/run && for i in 1 2 3 4 5; do timeout 300 $B/demo 0 100 | grep -E "Decompression time|mismatch" | tr '\n' ' '; echo; done)
  echo "=== redundancy $r: reference gap-array encoder + decoder ('dec,' = kernel ms, mean of 100 runs) ==="
  (cd $W && timeout 300 $R/oracle/_ref/gap_encoder data100_100.bin ref.huff 2>&1 | tail -4; timeout 600 $R/oracle/_ref/gap_decoder ref.huff decoded_ref 2>&1 | tail -5; cmp decodedfile data100_100.bin && echo "reference decoder output: identical" || echo "reference decoder output: DIFFERS")
  echo "=== redundancy $r: our bin/decoder on the reference's .huff ==="
  (cd $W && timeout 300 $B/decoder ref.huff decoded_ours 2>&1 | tail -3; cmp decoded_ours data100_100.bin && echo "our decoder output: identical")
  echo "=== redundancy $r: our bin/encoder --device (and the host producer), then bin/decoder on its file ==="
  (cd $W && timeout 300 $B/encoder --device data100_100.bin ours_dev.huff 2>&1 | tail -3; timeout 300 $B/encoder data100_100.bin ours_host.huff 2>&1 | tail -2; cmp ours_dev.huff ours_host.huff && echo "device-encoded .huff == host-encoded .huff"; timeout 300 $B/decoder ours_dev.huff decoded_dev 2>&1 | tail -3; cmp decoded_dev data100_100.bin && echo "round trip through the device encoder: identical")
done
rm -rf $W
