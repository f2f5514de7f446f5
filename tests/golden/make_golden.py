"""Generates tests/golden/llhuff_golden.npz in THIS container (needs /root/reference
built into oracle/_ref by oracle/Makefile).  For each case the input bytes come
from the seeded restatement of generate.cpp (oracle/gen_restatement.cpp) and the
stream + decode table from the UNMODIFIED reference encoder
(gpuhd/encoder/src/llhuffman_encoder.cc via oracle/ref_wrap/llhuff_ref.cpp); the
reference CPU codec (sequential.cpp) is run on the same bytes and its verdict
recorded.  The GPU box has no reference tree: tests there read only the .npz."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

CASES = [  # (name, n, redundancy, seed)
    ("r000", 3001, 0.0, 1), ("r010", 3001, 0.1, 2), ("r025", 3000, 0.25, 3), ("r050", 4099, 0.5, 4),
    ("r075", 3003, 0.75, 5), ("r090", 3007, 0.9, 6), ("r099", 3011, 0.99, 7), ("r100", 2999, 1.0, 8),
    ("tiny5", 5, 0.5, 9), ("tiny40", 40, 0.5, 10),
]


def d1_case():
    """Smallest n (r=0.5, seed 4) whose last codeword straddles a unit boundary: the reference packer
    drops its tail (defect D1, llhuffman_encoder.cc:213-237)."""
    for n in range(1000, 1100):
        d = O.generate(n, 0.5, 4)
        ref = O.ref_llhuff_encode(d)
        dec = O.decode_flat(ref["units"], ref["lut"], n)
        if not np.array_equal(dec, d) and np.array_equal(dec[:-1], d[:-1]):
            return n
    raise RuntimeError("no D1 case found")


def main():
    out = {}
    meta = {}
    cases = list(CASES) + [("d1", d1_case(), 0.5, 4)]
    for name, n, r, seed in cases:
        d = O.generate(n, r, seed)
        ref = O.ref_llhuff_encode(d)
        dec = O.decode_flat(ref["units"], ref["lut"], n)
        ok_seq, dec_seq, _, _, _ = O.ref_sequential_roundtrip(d)
        out[name + "_data"] = d
        out[name + "_units"] = ref["units"]
        out[name + "_lut"] = ref["lut"]
        meta[name] = dict(n=n, redundancy=r, seed=seed, sha256_data=hashlib.sha256(d.tobytes()).hexdigest(),
                          sha256_units=hashlib.sha256(ref["units"].tobytes()).hexdigest(),
                          oracle_equals_original=bool(np.array_equal(dec, d)),
                          oracle_equals_original_but_last=bool(np.array_equal(dec[:-1], d[:-1])),
                          reference_sequential_pass=bool(ok_seq and np.array_equal(dec_seq, d)),
                          max_len=int(ref["lens"].max()))
        print(name, meta[name])
    np.savez_compressed(os.path.join(HERE, "llhuff_golden.npz"), **out)
    # generator pin: hashes of larger seeded inputs (the generator must reproduce them anywhere)
    gen = {}
    for r, seed in [(0.0, 1), (0.5, 2), (0.9, 3), (1.0, 4)]:
        d = O.generate(1 << 20, r, seed)
        gen["r%03d_s%d" % (int(r * 100), seed)] = hashlib.sha256(d.tobytes()).hexdigest()
    with open(os.path.join(HERE, "llhuff_golden.json"), "w") as f:
        json.dump(dict(cases=meta, generator_1MiB=gen), f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
