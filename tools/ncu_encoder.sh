#!/bin/bash
# full ncu capture of the device encoder's kernels (histogram, code construction, pack) on 100 MiB, after a plain run
mkdir -p gpurun_out
python - <<PY
import sys; sys.path.insert(0, ".")
import hd_b200 as hd
hd.generate(100 << 20, 0.5, seed=1).tofile("/tmp/hd_enc_in.bin")
PY
B=cse375-finalproj-huffman-decoding_b200/bin
$B/encoder --device /tmp/hd_enc_in.bin /tmp/hd_enc_out.huff > gpurun_out/plain_enc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hd_ -o gpurun_out/r2_encoder_100m $B/encoder --device /tmp/hd_enc_in.bin /tmp/hd_enc_out.huff > gpurun_out/ncu_enc.log 2>&1
tail -3 gpurun_out/plain_enc.log
