"""CPU tests: the product's per-thread kernel source (csrc/hd_core.h), stepped by the emulator under
tests/emul with the kernels' tile/block structure, against the oracle.  Covers the same edge cases as the
GPU parity suite so that control-flow bugs are caught without a GPU."""
import numpy as np
import pytest

import _emul as E
import _oracle as O
import cases


def _nib(w, i):
    return (w[i >> 3] >> (4 * (i & 7))) & 15


def _check_both(data, pairs, phase=0):
    n = len(data)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    n_seg = (len(units) + 3) // 4
    # gap path on the oracle packer's gap array
    got = E.decode_gap(units, np.concatenate([gw, np.zeros(2, np.uint32)]), n_seg, n, pairs=pairs, phase=phase)
    assert np.array_equal(got, data)
    # self-sync: emulated K1a/K1b(/K1c) produce the gap array, then the same K4.  Both landing rules: the fast one
    # (whole LUT jumps: landings may lie past the first codeword of a segment) and the exact .huff one.
    g2, rounds, stats = E.sync(units, pairs=pairs)
    got2 = E.decode_gap(units, g2, n_seg, n, pairs=pairs, phase=phase)
    assert np.array_equal(got2, data)
    g3, _, _ = E.sync(units, pairs=pairs, exact=True)
    got3 = E.decode_gap(units, g3, n_seg, n, pairs=pairs, phase=phase)
    assert np.array_equal(got3, data)
    return rounds, stats


def test_golden_vectors_through_emulated_kernels():
    gold, _ = cases.golden()
    for name, c in gold.items():
        n = len(c["data"])
        units, lut = c["units"], c["lut"]
        n_seg = (len(units) + 3) // 4
        g2, rounds, stats = E.sync(units, flat_lut=lut)
        got = E.decode_gap(units, g2, n_seg, n, flat_lut=lut)
        want = O.decode_flat(units, lut, n)
        assert np.array_equal(got, want), name


@pytest.mark.parametrize("r", cases.SWEEP)
def test_sweep(r):
    n = 150001
    data = O.generate(n, r, seed=int(r * 100) + 1)
    syms, lengths = O.huffman_lengths(data, 11)
    _check_both(data, cases.pairs_of(syms, lengths), phase=int(r * 100) % 16)


@pytest.mark.parametrize("n", [1, 2, 3, 15, 16, 17, 31, 43, 127, 128, 129, 255, 256, 257, 1023, 4096])
def test_small_sizes(n):
    data = O.generate(n, 0.5, seed=n)
    syms, lengths = O.huffman_lengths(data, 11)
    _check_both(data, cases.pairs_of(syms, lengths), phase=n % 16)


def test_single_symbol():
    data = np.full(3000, 65, dtype=np.uint8)
    _check_both(data, cases.pairs_of([65], [1]))


def test_never_synchronising_code_needs_sequential_chain():
    pairs = cases.fixed_length_code(8, 3)
    data = cases.data_for_pairs(pairs, 200001, seed=5)
    rounds, stats = _check_both(data, pairs)
    assert stats[0] > 0 and stats[1] >= 3     # published exits moved, K1b rounds ran: the sequential-chain fallback


@pytest.mark.parametrize("length", [1, 2, 4, 5, 7, 8, 11])
def test_fixed_length_codes(length):
    k = min(256, 1 << length)
    pairs = cases.fixed_length_code(k, length)
    data = cases.data_for_pairs(pairs, 50003, seed=length)
    _check_both(data, pairs)


def test_long_codes_up_to_16_bits():
    pairs = cases.skewed_long_code()
    assert E.table_info(pairs=pairs)["has_long"] == 1
    for geometric in (True, False):
        data = cases.data_for_pairs(pairs, 60007, seed=9, geometric=geometric)
        _check_both(data, pairs)


def test_output_clamped_to_n_symbols():
    data = O.generate(20000, 0.9, seed=3)
    syms, lengths = O.huffman_lengths(data, 11)
    pairs = cases.pairs_of(syms, lengths)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    n_seg = (len(units) + 3) // 4
    for want_n in (20000, 19999, 10000, 17):
        got = E.decode_gap(units, np.concatenate([gw, np.zeros(2, np.uint32)]), n_seg, want_n, pairs=pairs)
        assert len(got) == want_n and np.array_equal(got, data[:want_n])


def test_sync_gap_array_equals_oracle_and_encoder_rule():
    for r in (0.25, 0.5, 0.9):
        data = O.generate(120000, r, seed=21)
        syms, lengths = O.huffman_lengths(data, 11)
        pairs = cases.pairs_of(syms, lengths)
        units, gw_enc, bits, lut = cases.oracle_stream(data, pairs)
        g2, _, _ = E.sync(units, pairs=pairs, exact=True)     # exact landings == .huff gap semantics
        off, cnt = O.sync_points(units, lut, len(data))
        n_real = int(np.nonzero(cnt)[0].max())
        idx = np.arange(n_real)
        want = O.pack_gap_words(off[1:])
        assert np.array_equal(_nib(g2, idx), _nib(want, idx))
        assert np.array_equal(_nib(g2, idx), _nib(gw_enc, idx))
