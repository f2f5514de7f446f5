"""Shared input builders for the CPU and GPU test suites."""
import json
import os

import numpy as np

import _oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_NPZ = os.path.join(HERE, "golden", "llhuff_golden.npz")
GOLDEN_JSON = os.path.join(HERE, "golden", "llhuff_golden.json")
SWEEP = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 1.0]   # redundancies of BASELINE.json configs[4]


def golden():
    z = np.load(GOLDEN_NPZ)
    with open(GOLDEN_JSON) as f:
        meta = json.load(f)
    cases = {}
    for name, m in meta["cases"].items():
        cases[name] = dict(data=z[name + "_data"], units=z[name + "_units"], lut=z[name + "_lut"], meta=m)
    return cases, meta


def pairs_of(syms, lengths):
    return np.stack([np.asarray(syms, dtype=np.uint8), np.asarray(lengths, dtype=np.uint8)], axis=1)


def fixed_length_code(n_symbols, length):
    """All codes `length` bits: a prefix code that NEVER self-synchronises when 128 % length != 0."""
    syms = np.arange(n_symbols, dtype=np.uint8)
    return pairs_of(syms, np.full(n_symbols, length, dtype=np.uint8))


def skewed_long_code():
    """Lengths 1,2,...,15,16,16: exercises codes longer than the 12-bit primary LUT (.huff allows 16)."""
    lengths = list(range(1, 16)) + [16, 16]
    return pairs_of(np.arange(len(lengths), dtype=np.uint8) + 40, np.array(lengths, dtype=np.uint8))


def data_for_pairs(pairs, n, seed, geometric=False):
    rng = np.random.default_rng(seed)
    k = len(pairs)
    if geometric:
        p = np.array([2.0 ** -int(l) for l in pairs[:, 1]])
        p /= p.sum()
        idx = rng.choice(k, size=n, p=p)
    else:
        idx = rng.integers(0, k, size=n)
    return pairs[idx, 0].astype(np.uint8)


def oracle_stream(data, pairs, pad_unit=False):
    """Pack with the ORACLE packer (independent of the product encoder)."""
    codes, lens = O.canonical_codes(pairs[:, 0], pairs[:, 1])
    units, bits, gaps = O.pack(data, codes, lens, flush_tail=True, want_gaps=True)
    gw = O.pack_gap_words(gaps)
    if pad_unit:
        units = np.concatenate([units, np.zeros(1, dtype=np.uint32)])
    maxlen = int(pairs[:, 1].max())
    lut = O.flat_lut(codes, lens, maxlen)
    return units, gw, bits, lut
