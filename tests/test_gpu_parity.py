"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, the committed golden vectors
produced by the reference encoder, and size-independent round-trip properties at BASELINE.json's sizes.
Bit-exact: this is byte/integer work, no tolerance."""
import hashlib

import numpy as np
import pytest

import _oracle as O
import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def _selfsync(hd, units, lut, n, first_bit=0):
    d_units = hd.to_device_units(units)
    tab = hd.Codetable(flat_lut=lut)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    out = dec.decode_selfsync(d_units, tab, n, first_bit=first_bit)
    res = out.cpu().numpy()
    cnt = int(dec.count.item())
    tab.close()
    return res, cnt


def _gap(hd, units, gap_words, pairs, n):
    import torch
    d_units = hd.to_device_units(units)
    d_gaps = torch.from_numpy(np.ascontiguousarray(gap_words, dtype=np.uint32).view(np.int32)).cuda()
    tab = hd.Codetable(pairs=pairs)
    dec = hd.Decoder(d_units.numel(), selfsync=False)
    out = dec.decode_gap(d_units, d_gaps, tab, n)
    res = out.cpu().numpy()
    tab.close()
    return res


# ---------------------------------------------------------------- golden vectors (reference encoder output)
def test_selfsync_golden_vectors(hd, torch_cuda):
    gold, _ = cases.golden()
    for name, c in gold.items():
        n = len(c["data"])
        got, cnt = _selfsync(hd, c["units"], c["lut"], n)
        want = O.decode_flat(c["units"], c["lut"], n)
        assert np.array_equal(got, want), name
        assert cnt == n, name
        if c["meta"]["oracle_equals_original"]:
            assert np.array_equal(got, c["data"]), name
        else:   # reference encoder defect D1: last symbol lost by the ENCODER
            assert np.array_equal(got[:-1], c["data"][:-1]), name


# ---------------------------------------------------------------- redundancy sweep, both schemes
@pytest.mark.parametrize("r", cases.SWEEP)
@pytest.mark.parametrize("n", [1 << 20, (1 << 20) + 4097])
def test_selfsync_sweep(hd, torch_cuda, r, n):
    data = O.generate(n, r, seed=int(r * 100) + 3)
    pairs = hd.build_code(data, hd.MAX_CODEWORD_LENGTH_CUHD)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    got, cnt = _selfsync(hd, units, lut, n)
    want = O.decode_flat(units, lut, n)
    assert np.array_equal(got, want)
    assert np.array_equal(got, data)
    assert cnt == n


@pytest.mark.parametrize("r", cases.SWEEP)
def test_gap_sweep(hd, torch_cuda, r):
    n = (1 << 20) + 12345
    data = O.generate(n, r, seed=int(r * 100) + 7)
    pairs = hd.build_code(data, hd.MAX_CODEWORD_LENGTH_GAP)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    image = hd.huff_write(pairs, n, units, gaps)
    want = O.huff_decode(image, n)
    got, _ = hd.decode_huff(image, device=0)
    assert np.array_equal(got, want)
    assert np.array_equal(got, data)
    got2 = _gap(hd, units, gaps, pairs, n)
    assert np.array_equal(got2, data)


def test_gap_array_from_selfsync_matches_oracle(hd, torch_cuda):
    for r in (0.25, 0.5, 0.9):
        n = 600000
        data = O.generate(n, r, seed=21)
        pairs = hd.build_code(data, 11)
        lut = hd.flat_lut(pairs, 11)
        units, gaps_enc, bits = hd.encode(data, pairs, want_gaps=True)
        d_units = hd.to_device_units(units)
        tab = hd.Codetable(flat_lut=lut)
        dec = hd.Decoder(d_units.numel(), selfsync=True)
        gw = dec.gap_array(d_units, tab).cpu().numpy().view(np.uint32)
        off, cnt = O.sync_points(units, lut, n)
        n_real = int(np.nonzero(cnt)[0].max())     # segments 1..n_real hold codeword starts
        want = O.pack_gap_words(off[1:])
        nib = lambda w, i: (w[i >> 3] >> (4 * (i & 7))) & 15
        idx = np.arange(n_real)
        assert np.array_equal(nib(gw, idx), nib(want, idx))
        # and equals what the ENCODER recorded (encoder.cu:307-312 rule)
        assert np.array_equal(nib(gw, idx), nib(gaps_enc, idx))
        tab.close()


# ---------------------------------------------------------------- edge cases
@pytest.mark.parametrize("n", [1, 2, 3, 15, 16, 17, 31, 43, 127, 128, 129, 255, 256, 257, 1023, 4096, 65537])
def test_small_sizes_both_schemes(hd, torch_cuda, n):
    data = O.generate(n, 0.5, seed=n)
    pairs = hd.build_code(data, 11)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    got, cnt = _selfsync(hd, units, lut, n)
    assert np.array_equal(got, data)
    units2, gw2, bits2, _ = cases.oracle_stream(data, pairs)
    got2 = _gap(hd, units2, gw2, pairs, n)
    assert np.array_equal(got2, data)


def test_single_symbol_input(hd, torch_cuda):
    data = np.full(5000, 65, dtype=np.uint8)
    pairs = hd.build_code(data, 11)
    assert pairs.tolist() == [[65, 1]]
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    got, _ = _selfsync(hd, units, lut, len(data))
    assert np.array_equal(got, data)
    got2 = _gap(hd, units[:-1], gw, pairs, len(data))
    assert np.array_equal(got2, data)


def test_never_synchronising_code(hd, torch_cuda):
    """All-3-bit code: misaligned parses never merge, so the fixed point degrades to a sequential chain
    (the reference emits garbage here: SURVEY D3/D4).  Result must still be exact."""
    pairs = cases.fixed_length_code(8, 3)
    n = 700001
    data = cases.data_for_pairs(pairs, n, seed=5)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    got, cnt = _selfsync(hd, units, lut, n)
    assert np.array_equal(got, data)
    assert cnt == n


@pytest.mark.parametrize("length,n", [(3, 60_000_017), (5, 20_000_003), (7, 12_345_678)])
def test_never_synchronising_code_at_size(hd, torch_cuda, length, n):
    """The same at sizes with thousands of K1a tiles (60 M symbols of a 3-bit code = 2800 tiles of 64 Kbit): after a few
    K1b rounds the synchronisation switches to the transfer-function pass (every tile's entry -> exit map in parallel,
    one chase through the maps, K1a once more with exact tile entries) instead of one round per tile; also the EXACT
    variant behind hd_selfsync_gap_array, whose result must be the gap array of the stream."""
    import time
    torch = torch_cuda
    k = min(256, 1 << length)
    pairs = cases.fixed_length_code(k, length)
    data = hd.generate(n, 0.0, seed=length) % k
    data = data.astype(np.uint8)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    units = np.concatenate([units, np.zeros(1, dtype=np.uint32)])          # the CUHD pad unit
    lut = hd.flat_lut(pairs, 11)
    tab = hd.Codetable(flat_lut=lut)
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    dec.decode_selfsync(d_units, tab, n)                                   # (kernels loaded)
    torch.cuda.synchronize()
    t0 = time.time()
    out = dec.decode_selfsync(d_units, tab, n)
    torch.cuda.synchronize()
    dt = time.time() - t0
    assert bool(torch.equal(out, torch.from_numpy(data).cuda())) and int(dec.count.item()) == n
    assert dt < 2.0, "never-synchronising decode took %.2f s" % dt
    gw = dec.gap_array(d_units, tab).cpu().numpy().view(np.uint32)
    n_seg = (len(units) - 1 + 3) // 4
    assert np.array_equal(gw[:(n_seg + 7) // 8 - 1], gaps[:(n_seg + 7) // 8 - 1])   # (the last word also covers the pad unit)
    tab.close()


@pytest.mark.parametrize("length", [1, 2, 4, 5, 7, 8, 11])
def test_fixed_length_codes(hd, torch_cuda, length):
    k = min(256, 1 << length)
    pairs = cases.fixed_length_code(k, length)
    n = 200003
    data = cases.data_for_pairs(pairs, n, seed=length)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    got, _ = _selfsync(hd, units, lut, n)
    assert np.array_equal(got, data)
    got2 = _gap(hd, units[:-1], gw, pairs, n)
    assert np.array_equal(got2, data)


@pytest.mark.parametrize("length", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [1, 17, 4095, 200003])
def test_fixed_length_fast_path_equals_general_kernels(hd, torch_cuda, length, n):
    """Codes whose codewords all have the same length L | 8 take the dedicated kernel (hd_fixed.cuh) when the stream
    starts at bit 0 and the output is 16-byte aligned; hd_set_tuning(-1, 0) keeps them on K1a + K4.  Both must give the
    original bytes, the same symbol count, and write nothing past the requested number of symbols."""
    torch = torch_cuda
    pairs = cases.fixed_length_code(1 << length, length)
    data = cases.data_for_pairs(pairs, n, seed=100 + length)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    d_units = hd.to_device_units(units)
    d_units_gap = hd.to_device_units(units[:-1])
    d_gaps = torch.from_numpy(np.ascontiguousarray(gw, dtype=np.uint32).view(np.int32)).cuda()
    tab_ss = hd.Codetable(flat_lut=lut)
    tab_gap = hd.Codetable(pairs=pairs)
    dec_ss = hd.Decoder(d_units.numel(), selfsync=True)
    dec_gap = hd.Decoder(d_units_gap.numel(), selfsync=False)
    try:
        for fast in (True, False):
            hd.set_tuning(0 if fast else -1, 0)
            launches = hd.launch_count()
            for want_n in sorted({n, max(1, n - 1), max(1, n // 2)}):
                for shift in (0, 3):   # an unaligned output always takes the general kernels
                    buf = torch.full((n + 64,), 0xEE, dtype=torch.uint8, device="cuda")
                    dec_ss.decode_selfsync(d_units, tab_ss, want_n, out=buf[shift:shift + want_n])
                    res = buf.cpu().numpy()
                    assert np.array_equal(res[shift:shift + want_n], data[:want_n]), ("selfsync", fast, want_n, shift)
                    assert (res[:shift] == 0xEE).all() and (res[shift + want_n:] == 0xEE).all()
                    assert int(dec_ss.count.item()) == want_n
                    buf.fill_(0xEE)
                    dec_gap.decode_gap(d_units_gap, d_gaps, tab_gap, want_n, out=buf[shift:shift + want_n])
                    res = buf.cpu().numpy()
                    assert np.array_equal(res[shift:shift + want_n], data[:want_n]), ("gap", fast, want_n, shift)
                    assert (res[:shift] == 0xEE).all() and (res[shift + want_n:] == 0xEE).all()
            del launches
    finally:
        hd.set_tuning(0, 0)
        tab_ss.close()
        tab_gap.close()


def test_long_codes_up_to_16_bits(hd, torch_cuda):
    pairs = cases.skewed_long_code()
    for geometric in (True, False):
        n = 300007
        data = cases.data_for_pairs(pairs, n, seed=9, geometric=geometric)
        units, gw, bits, lut = cases.oracle_stream(data, pairs)
        image = hd.huff_write(pairs, n, units, gw)
        want = O.huff_decode(image, n)
        assert np.array_equal(want, data)
        got, _ = hd.decode_huff(image, device=0)
        assert np.array_equal(got, data)
        # self-sync on the same 16-bit code (flat LUT of 2^16 entries)
        units_p = np.concatenate([units, np.zeros(1, dtype=np.uint32)])
        got2, _ = _selfsync(hd, units_p, lut, n)
        assert np.array_equal(got2, data)


def test_max_length_exactly_11(hd, torch_cuda):
    gold, _ = cases.golden()
    c = gold["r050"]
    assert c["meta"]["max_len"] == 11
    got, _ = _selfsync(hd, c["units"], c["lut"], len(c["data"]))
    assert np.array_equal(got, c["data"])


def test_unaligned_output_pointer(hd, torch_cuda):
    torch = torch_cuda
    n = 300000
    data = O.generate(n, 0.5, seed=77)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    d_units = hd.to_device_units(units)
    tab = hd.Codetable(flat_lut=lut)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    for shift in (1, 3, 8, 15):
        buf = torch.full((n + 64,), 0xAB, dtype=torch.uint8, device="cuda")
        out = buf[shift:shift + n]
        dec.decode_selfsync(d_units, tab, n, out=out)
        res = buf.cpu().numpy()
        assert np.array_equal(res[shift:shift + n], data)
        assert (res[:shift] == 0xAB).all() and (res[shift + n:] == 0xAB).all()   # no OOB writes
    tab.close()


def test_output_clamped_to_n_symbols(hd, torch_cuda):
    torch = torch_cuda
    n = 100000
    data = O.generate(n, 0.9, seed=3)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    d_units = hd.to_device_units(units)
    tab = hd.Codetable(flat_lut=lut)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    for want_n in (n, n - 1, n // 2, 17):
        buf = torch.full((n + 32,), 0xCD, dtype=torch.uint8, device="cuda")
        dec.decode_selfsync(d_units, tab, want_n, out=buf[:want_n])
        res = buf.cpu().numpy()
        assert np.array_equal(res[:want_n], data[:want_n])
        assert (res[want_n:] == 0xCD).all()
    tab.close()


def test_repeated_decodes_reuse_workspace(hd, torch_cuda):
    n = 500000
    data = O.generate(n, 0.5, seed=13)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    d_units = hd.to_device_units(units)
    tab = hd.Codetable(flat_lut=lut)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    for _ in range(5):
        out = dec.decode_selfsync(d_units, tab, n)
        assert np.array_equal(out.cpu().numpy(), data)
    tab.close()


def test_sharded_gap_decode_concatenates(hd, torch_cuda):
    """Multi-GPU partitioning logic on one device: shards on tile boundaries, first_bit from the gap array,
    per-shard counts -> exclusive scan -> concatenation equals the whole."""
    import torch
    n = 3000000
    data = O.generate(n, 0.5, seed=31)
    pairs = hd.build_code(data, 16)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    n_seg = (len(units) + 3) // 4
    tab = hd.Codetable(pairs=pairs)
    for world in (2, 4, 8):
        parts = []
        for rank in range(world):
            b, e = hd.shard_segments(n_seg, rank, world)
            assert b % 256 == 0
            su = units[4 * b:min(4 * e + 1, len(units))]     # +1 unit: the shard's last codeword may end there
            sg = gaps[b // 8:]
            fb = 0 if b == 0 else int((gaps[(b - 1) >> 3] >> (4 * ((b - 1) & 7))) & 15)
            d_units = hd.to_device_units(su)
            d_gaps = torch.from_numpy(np.ascontiguousarray(sg).view(np.int32)).cuda()
            dec = hd.Decoder(d_units.numel(), selfsync=False)
            out = dec.decode_gap(d_units, d_gaps, tab, n, first_bit=fb,    # capacity n, count tells the truth
                                 n_segments=min(e, n_seg) - b)
            cnt = int(dec.count.item())
            parts.append(out[:cnt].cpu().numpy())
        whole = np.concatenate(parts)
        # the last shard also decodes the pad bits of the final unit; the global clamp to n_symbols happens
        # after the scan of the per-shard counts (each shard alone cannot know how many symbols precede it)
        assert n <= len(whole) <= n + 31 and np.array_equal(whole[:n], data)
    tab.close()


# ---------------------------------------------------------------- BASELINE.json sizes: round-trip properties
def _roundtrip_large(hd, torch, n, r, scheme):
    data = hd.generate(n, r, seed=1234)
    if scheme == "selfsync":
        pairs = hd.build_code(data, 11)
        units, bits = hd.encode(data, pairs, pad_unit=True)
        tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
        d_units = hd.to_device_units(units)
        dec = hd.Decoder(d_units.numel(), selfsync=True)
        out = dec.decode_selfsync(d_units, tab, n)
    else:
        pairs = hd.build_code(data, 16)
        units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
        tab = hd.Codetable(pairs=pairs)
        d_units = hd.to_device_units(units)
        d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
        dec = hd.Decoder(d_units.numel(), selfsync=False)
        out = dec.decode_gap(d_units, d_gaps, tab, n)
    ref = torch.from_numpy(data).cuda()
    same = bool(torch.equal(out, ref))
    cnt = int(dec.count.item())
    tab.close()
    return same, cnt


def test_100MiB_selfsync_roundtrip(hd, torch_cuda):
    """configs[1]: gpuhd self-sync decode, 100 MB generate.cpp input (./bin/demo 0 100 shape)."""
    n = 100 * 1024 * 1024
    same, cnt = _roundtrip_large(hd, torch_cuda, n, 0.5, "selfsync")
    assert same and cnt == n


def test_1GiB_gap_roundtrip(hd, torch_cuda):
    """configs[2]: gap-array decode of a .huff-layout stream, 1 GB input."""
    n = 1 << 30
    same, cnt = _roundtrip_large(hd, torch_cuda, n, 0.5, "gap")
    assert same and cnt == n


# ---------------------------------------------------------------- multi-GPU layer (hd_mgpu_*), ranks emulated on one device
class _ThreadGather:
    """A real collective between `world` host threads of this process (one thread per emulated rank, all on cuda:0):
    the hd_allgather_fn callback handed to hd_mgpu_exchange_callback."""

    def __init__(self, world, torch):
        import threading
        self.world, self.torch = world, torch
        self.barrier = threading.Barrier(world)
        self.table = np.zeros((world, 3), dtype=np.uint64)

    def callback(self, rank):
        torch, world = self.torch, self.world

        class Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}

        def allgather(d_send, d_recv, nbytes, stream, user):
            try:
                torch.cuda.synchronize()
                send = torch.as_tensor(Dev(d_send, 3), device="cuda")
                self.table[rank] = send.cpu().numpy().view(np.uint64)
                self.barrier.wait(timeout=120)
                recv = torch.as_tensor(Dev(d_recv, 3 * world), device="cuda")
                recv.copy_(torch.from_numpy(self.table.reshape(-1).view(np.int64).copy()))
                torch.cuda.synchronize()
                self.barrier.wait(timeout=120)
                return 0
            except Exception:
                return 1
        return allgather


def _run_ranks(torch, world, make):
    """make(rank, allgather) -> a sharded decoder; every emulated rank runs step() + finish() on its own host thread.
    Returns (whole output, first-round infos of rank 0, placement, repairs per rank)."""
    import threading
    tg = _ThreadGather(world, torch)
    res, err = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(0)
            d = make(rank, tg.callback(rank) if world > 1 else None)
            d.step()
            placed = d.finish()
            res[rank] = (d.out[:d.count_local].cpu().numpy(), d.infos().copy(), placed, d.repairs(), d.offset)
            tg.barrier.wait(timeout=120) if world > 1 else None
            d.dist = None
            d.close()
        except Exception as e:   # noqa: BLE001
            err.append(e)
            tg.barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if err:
        raise err[0]
    for r in range(1, world):
        assert np.array_equal(res[r][1], res[0][1]) and res[r][2] == res[0][2]     # every rank holds the same triples
    return np.concatenate([x[0] for x in res]), res[0][1], res[0][2], [x[3] for x in res]


def _sharded_selfsync(hd, torch, units, lut, n, world, halo_segments):
    tab = hd.Codetable(flat_lut=lut)
    out = _run_ranks(torch, world, lambda r, ag: hd.ShardedSelfSyncDecoder(units, tab, n, r, world, dist=None,
                                                                           halo_segments=halo_segments, allgather=ag))
    tab.close()
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_selfsync_shards_with_halo(hd, torch_cuda, world):
    n = 5_000_000
    data = O.generate(n, 0.5, seed=41)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    whole, infos, placed, repairs = _sharded_selfsync(hd, torch_cuda, units, lut, n, world, 4096)
    assert repairs == [0] * world                        # the halo synchronised every seam
    assert hd.seams_ok(infos) == []
    assert sum(c for _, c in placed) == n
    assert np.array_equal(whole, data)
    # every entry offset is a codeword boundary of the oracle's true parse at or after the first one of its segment
    # (the fast synchronisation records where a whole LUT jump lands, see sync_segment in csrc/hd_core.h)
    off, cnt = O.sync_points(units, lut, n)
    plan = hd.plan_shards(len(units), world)
    codes, lens = O.canonical_codes(pairs[:, 0], pairs[:, 1])
    for g in range(1, world):
        seg = plan[g][0]
        entry = int(infos[g][1]) & 0xffff
        assert int(off[seg]) <= entry <= 15
        # walk the true parse from the segment's first codeword: the entry must be one of its boundaries
        k = int(cnt[:seg].sum())
        pos = int(off[seg])
        while pos < entry:
            pos += int(lens[data[k]])
            k += 1
        assert pos == entry


def test_selfsync_shards_repair_when_code_never_synchronises(hd, torch_cuda):
    """All-3-bit code: the halo cannot synchronise, hd_mgpu_finish catches it at the seams and the shards are decoded
    again from their predecessors' exact exit offsets (sequential chain across ranks)."""
    pairs = cases.fixed_length_code(8, 3)
    n = 2_000_003
    data = cases.data_for_pairs(pairs, n, seed=6)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    whole, infos, placed, repairs = _sharded_selfsync(hd, torch_cuda, units, lut, n, 4, 1024)
    assert np.array_equal(whole, data)
    assert sum(c for _, c in placed) == n
    assert sum(repairs[1:]) >= 3                       # every later shard had to be decoded again at least once


def test_selfsync_shards_more_ranks_than_tiles(hd, torch_cuda):
    """fewer 256-segment tiles than ranks -> ranks without segments publish an empty triple instead of
    failing (which used to leave the other ranks waiting in the exchange)."""
    n = 9000                                           # ~1670 units = 2 tiles
    data = O.generate(n, 0.5, seed=43)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    whole, infos, placed, repairs = _sharded_selfsync(hd, torch_cuda, units, lut, n, 4, 4096)
    assert np.array_equal(whole, data) and sum(c for _, c in placed) == n
    assert sum(1 for row in infos if int(row[1]) == (1 << 64) - 1) >= 2


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("r", [0.5, 0.9])
def test_gap_shards_through_hd_mgpu(hd, torch_cuda, world, r):
    """gap-array shards (segment entry rule decoder.cu:505-507) through hd_mgpu_* -- no halo, no
    repair, only the counts place the outputs.  16-bit limited .huff code."""
    n = 6_000_011
    data = O.generate(n, r, seed=int(100 * r) + world)
    pairs = hd.build_code(data, 16)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    tab = hd.Codetable(pairs=pairs)
    whole, infos, placed, repairs = _run_ranks(
        torch_cuda, world, lambda rk, ag: hd.ShardedGapDecoder(units, gaps, tab, n, rk, world, dist=None, allgather=ag))
    tab.close()
    assert repairs == [0] * world and hd.seams_ok(infos) == []
    assert sum(c for _, c in placed) == n and np.array_equal(whole, data)
    want = O.decode_gap(units, gaps, (len(units) + 3) // 4, O.flat_lut(*O.canonical_codes(pairs[:, 0], pairs[:, 1]), 16), n)
    assert np.array_equal(whole, want)


def test_host_decoder_shard_entries(hd, torch_cuda):
    """hd_host_decode_selfsync_shard / hd_host_decode_gap_shard (the per-rank pipeline of hd_mgpu_host_step) on the
    middle shard of a three-way split: bytes == the matching slice of the original, triple == what the device path
    reports for the same shard."""
    import ctypes as C
    torch = torch_cuda
    L = hd.lib()
    n = 30_000_019
    data = O.generate(n, 0.5, seed=52)
    for scheme in (hd.SCHEME_SELFSYNC, hd.SCHEME_GAP):
        if scheme == hd.SCHEME_SELFSYNC:
            pairs = hd.build_code(data, 11)
            units, _ = hd.encode(data, pairs, pad_unit=True)
            tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
            gaps = None
        else:
            pairs = hd.build_code(data, 16)
            units, gaps, _ = hd.encode(data, pairs, want_gaps=True)
            tab = hd.Codetable(pairs=pairs)
        # reference: the device path of the same layer, three ranks emulated by threads
        make = ((lambda r, ag: hd.ShardedSelfSyncDecoder(units, tab, n, r, 3, dist=None, allgather=ag)) if gaps is None else
                (lambda r, ag: hd.ShardedGapDecoder(units, gaps, tab, n, r, 3, dist=None, allgather=ag)))
        whole, infos, placed, repairs = _run_ranks(torch, 3, make)
        assert np.array_equal(whole, data)
        p = hd.plan_shard(scheme, len(units), 1, 3, max(1, tab.min_length))
        h_units = torch.from_numpy(np.ascontiguousarray(units[p.first_unit:p.first_unit + p.readable_units]).view(np.int32)).pin_memory()
        cap = int(p.out_capacity)
        h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
        hdec = hd.HostDecoder(int(p.readable_units), cap, device=0)
        info = (C.c_uint64 * 3)()
        if gaps is None:
            st = L.hd_host_decode_selfsync_shard(hdec.handle, h_units.data_ptr(), int(p.halo_units), int(p.shard_units),
                                                 int(p.readable_units), tab.handle, h_out.data_ptr(), cap, info, None)
        else:
            gw = np.ascontiguousarray(gaps[p.first_gap_word:p.first_gap_word + p.n_gap_words])
            st = L.hd_host_decode_gap_shard(hdec.handle, h_units.data_ptr(), int(p.readable_units), gw[1:].ctypes.data,
                                            int(p.seg_end - p.seg_begin), int(gw[0] >> 28), tab.handle, h_out.data_ptr(),
                                            cap, info, None)
        assert st == 0
        off, cnt = placed[1]
        assert int(info[0]) == int(infos[1][0]) == cnt
        assert int(info[1]) == int(infos[1][1]) & 0xffff and int(info[2]) == int(infos[1][2])
        assert np.array_equal(h_out.numpy()[:cnt], data[off:off + cnt])
        hdec.close()
        tab.close()


def test_warmup_entry(hd, torch_cuda):
    assert hd.lib().hd_warmup() == 0


def test_host_decoder_selfsync_and_huff(hd, torch_cuda):
    torch = torch_cuda
    n = 3_000_001
    data = O.generate(n, 0.5, seed=51)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    hdec = hd.HostDecoder(len(units), n, device=0)
    h_units = torch.from_numpy(units.view(np.int32)).pin_memory()
    h_out = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        t = hdec.decode_selfsync(h_units, lut, h_out, n)
        assert np.array_equal(h_out.numpy(), data) and all(x >= 0 for x in t)
    pairs16 = hd.build_code(data, 16)
    units16, gaps, _ = hd.encode(data, pairs16, want_gaps=True)
    image = hd.huff_write(pairs16, n, units16, gaps)
    out2 = np.zeros(n, dtype=np.uint8)
    size, t = hdec.decode_huff(image, out2)
    assert size == n and np.array_equal(out2, data)
    hdec.close()


def test_host_decoder_falls_back_for_a_code_that_never_synchronises(hd, torch_cuda):
    """Pipelined host decode of a stream that spans several pieces with a fixed-length 3-bit code: no piece converges
    blind and no halo synchronises, so the chained fast path must hand over to the serial one (exact entry offsets)."""
    torch = torch_cuda
    pairs = cases.fixed_length_code(8, 3)
    n = 8_000_003                                # 187 k segments: more than the first (128 k segment) piece
    data = cases.data_for_pairs(pairs, n, seed=16)
    units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
    hdec = hd.HostDecoder(len(units), n, device=0)
    h_units = torch.from_numpy(units.view(np.int32)).pin_memory()
    h_out = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        h_out.zero_()
        hdec.decode_selfsync(h_units, lut, h_out, n)
        assert np.array_equal(h_out.numpy(), data)
    hdec.close()


def test_profile_hooks_report_kernel_times(hd, torch_cuda):
    n = 2_000_000
    data = O.generate(n, 0.5, seed=61)
    pairs = hd.build_code(data, 11)
    units, bits = hd.encode(data, pairs, pad_unit=True)
    tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    before = hd.launch_count()
    hd.profile_enable(True)
    out = dec.decode_selfsync(d_units, tab, n)
    ms = hd.profile_read()
    hd.profile_enable(False)
    assert hd.launch_count() - before >= 2
    assert ms["decode_write"] > 0 and ms["sync_chunks"] > 0
    assert np.array_equal(out.cpu().numpy(), data)
    tab.close()


def test_selfsync_enqueue_finish_split(hd, torch_cuda):
    """hd_decode_selfsync_enqueue + _finish == hd_decode_selfsync, also for a code that needs the fallback rounds."""
    torch = torch_cuda
    for pairs, data in ((None, O.generate(1_500_000, 0.5, seed=71)),
                        (cases.fixed_length_code(8, 3), None)):
        if pairs is None:
            pairs = hd.build_code(data, 11)
            units, bits = hd.encode(data, pairs, pad_unit=True)
            lut = hd.flat_lut(pairs, 11)
        else:
            data = cases.data_for_pairs(pairs, 400_001, seed=8)
            units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
        n = len(data)
        d_units = hd.to_device_units(units)
        tab = hd.Codetable(flat_lut=lut)
        dec = hd.Decoder(d_units.numel(), selfsync=True)
        out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            dec.decode_selfsync(d_units, tab, n, out=out, defer=True)
            dec.finish()
            assert np.array_equal(out.cpu().numpy(), data)
            assert int(dec.count.item()) == n
        tab.close()


@pytest.mark.parametrize("r", [0.0, 0.9, 1.0])
def test_1GiB_redundancy_sweep_roundtrip(hd, torch_cuda, r):
    """configs[4]: redundancy sweep at 1 GB (near-incompressible ... highly skewed; r=0.5 is covered above).
    Self-sync scheme with the llhuff 11-bit limit, byte-exact against the original."""
    n = 1 << 30
    same, cnt = _roundtrip_large(hd, torch_cuda, n, r, "selfsync")
    assert same and cnt == n


def test_device_built_tables_equal_host_built(hd, torch_cuda):
    """the table kernel (hd_table_from_lengths_device) writes, byte for byte, the blob that the host
    builder uploads -- for generate.cpp codes of every redundancy, fixed-length codes, 16-bit codes, one symbol."""
    torch = torch_cuda
    lists = []
    for r in cases.SWEEP:
        data = O.generate(200_000, r, seed=int(r * 100) + 9)
        lists.append(hd.build_code(data, 16))
        lists.append(hd.build_code(data, 11))
    lists += [cases.fixed_length_code(8, 3), cases.fixed_length_code(256, 8), cases.fixed_length_code(2, 1),
              cases.skewed_long_code(), cases.pairs_of([65], [1]), cases.pairs_of([1, 2, 3], [2, 2, 2])]
    for pairs in lists:
        host = hd.Codetable(pairs=pairs)
        dev = hd.Codetable(device_pairs=torch.from_numpy(np.ascontiguousarray(pairs)).cuda())
        assert (dev.min_length, dev.max_length) == (host.min_length, host.max_length)
        assert np.array_equal(dev.blob(), host.blob())
        host.close()
        dev.close()
    # malformed lists are refused: decreasing lengths, duplicate symbol, Kraft overflow
    for bad in (cases.pairs_of([1, 2], [3, 2]), cases.pairs_of([7, 7], [1, 1]), cases.pairs_of([1, 2, 3], [1, 1, 1])):
        with pytest.raises(hd.HDError):
            hd.Codetable(device_pairs=torch.from_numpy(np.ascontiguousarray(bad)).cuda())
    # and a decode through a device-built table
    n = 500_000
    data = O.generate(n, 0.5, seed=91)
    pairs = hd.build_code(data, 16)
    units, gaps, bits = hd.encode(data, pairs, want_gaps=True)
    tab = hd.Codetable(device_pairs=torch.from_numpy(np.ascontiguousarray(pairs)).cuda())
    d_units = hd.to_device_units(units)
    d_gaps = torch.from_numpy(gaps.view(np.int32)).cuda()
    dec = hd.Decoder(d_units.numel(), selfsync=False)
    assert np.array_equal(dec.decode_gap(d_units, d_gaps, tab, n).cpu().numpy(), data)
    tab.close()


@pytest.mark.parametrize("r", [0.0, 0.5, 0.9, 1.0])
@pytest.mark.parametrize("n", [1, 15, 16, 4095, 4097, 1_000_003])
def test_device_encoder_equals_host_encoder(hd, torch_cuda, r, n):
    """histogram + pack + gap array on the GPU give, bit for bit, the host producer's payload
    (itself checked against the oracle packer on CPU), and the payload decodes back."""
    torch = torch_cuda
    data = O.generate(n, r, seed=int(r * 100) + n % 97)
    d_data = torch.from_numpy(data).cuda()
    pairs, d_units, d_gaps, bits = hd.encode_device(d_data, 16)
    pairs_h = hd.build_code(data, 16)
    assert np.array_equal(pairs, pairs_h)
    units_h, gaps_h, bits_h = hd.encode(data, pairs_h, want_gaps=True)
    assert bits == bits_h
    assert np.array_equal(d_units.cpu().numpy().view(np.uint32), units_h)
    assert np.array_equal(d_gaps.cpu().numpy().view(np.uint32), gaps_h)
    tab = hd.Codetable(pairs=pairs)
    dec = hd.Decoder(max(d_units.numel(), 1), selfsync=False)
    out = dec.decode_gap(d_units, d_gaps, tab, n)
    assert np.array_equal(out.cpu().numpy(), data)
    tab.close()


@pytest.mark.parametrize("r,n,max_len", [(0.5, 1_000_003, 16), (0.9, 2_000_000, 16), (0.0, 300_001, 16), (1.0, 77_777, 16),
                                         (0.5, 500_000, 11), (0.25, 65_537, 9), (0.5, 300, 16)])
def test_device_code_construction_equals_host(hd, torch_cuda, r, n, max_len):
    """SURVEY 8(f) row 2, code construction off the host: hd_encode_device_auto (histogram -> package-merge KERNEL ->
    pack) gives the header list of hd_build_code (host package-merge, itself checked against the oracle's lengths) and
    the host packer's payload, bit for bit, and the payload decodes back."""
    torch = torch_cuda
    data = O.generate(n, r, seed=int(r * 100) + n % 89)
    pairs, d_units, d_gaps, bits = hd.encode_device_auto(torch.from_numpy(data).cuda(), max_len)
    pairs_h = hd.build_code(data, max_len)
    assert np.array_equal(pairs, pairs_h)
    units_h, gaps_h, bits_h = hd.encode(data, pairs_h, want_gaps=True)
    assert bits == bits_h
    assert np.array_equal(d_units.cpu().numpy().view(np.uint32), units_h)
    assert np.array_equal(d_gaps.cpu().numpy().view(np.uint32), gaps_h)
    tab = hd.Codetable(pairs=pairs)
    dec = hd.Decoder(max(d_units.numel(), 1), selfsync=False)
    out = dec.decode_gap(d_units.contiguous(), d_gaps.contiguous(), tab, n)
    assert np.array_equal(out.cpu().numpy(), data)
    tab.close()


def test_mailbox_single_rank_roundtrip(hd, torch_cuda):
    """hd_mailbox_* with world 1 (the multi-rank path runs under torchrun in bench.py): publish + wait move the triple
    into the gathered array, sequence numbers gate the wait."""
    import ctypes as C
    torch = torch_cuda
    L = hd.lib()
    h = C.c_void_p(None)
    assert L.hd_mailbox_create(1, 0, C.byref(h)) == 0
    handle = np.zeros(64, dtype=np.uint8)
    assert L.hd_mailbox_export(h, handle.ctypes.data) == 0
    assert L.hd_mailbox_connect(h, handle.ctypes.data) == 0
    info = torch.tensor([123456789012, 7, 9], dtype=torch.int64, device="cuda")
    allinfo = torch.zeros(3, dtype=torch.int64, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for seq in (1, 2, 3):
        info[0] += 1
        assert L.hd_mailbox_publish(h, C.c_void_p(info.data_ptr()), seq, sp) == 0
        assert L.hd_mailbox_wait(h, seq, C.c_void_p(allinfo.data_ptr()), sp) == 0
        torch.cuda.synchronize()
        assert allinfo.tolist() == info.tolist()
    assert L.hd_mailbox_timed_out(h) == 0
    assert L.hd_mailbox_publish(h, C.c_void_p(info.data_ptr()), 0, sp) != 0     # seq 0 is reserved
    L.hd_mailbox_destroy(h)


def test_beyond_4_gib(hd, torch_cuda):
    """4.5 GiB round trip entirely on the device (device encoder -> both decoders -> compare): 64-bit sizes end to end.
    The reference's formats and kernels stop at 4 GiB / 2 GiB (32-bit sizes); the property checked is the
    size-independent one, decode(encode(x)) == x."""
    torch = torch_cuda
    free, _ = torch.cuda.mem_get_info()
    if free < (24 << 30):
        pytest.skip("needs 24 GiB of free device memory")
    n = int(4.5 * (1 << 30)) + 12345
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    data = torch.empty(n, dtype=torch.uint8, device="cuda")
    step = 1 << 28
    for o in range(0, n, step):      # half of the mass on 4 symbols (the generate.cpp shape), the rest uniform
        m = min(step, n - o)
        u = torch.randint(0, 256, (m,), dtype=torch.int16, device="cuda", generator=g)
        hot = torch.randint(0, 2, (m,), dtype=torch.int16, device="cuda", generator=g)
        data[o:o + m] = torch.where(hot == 1, u & 3, u).to(torch.uint8)
        del u, hot
    for max_len, selfsync in ((16, False), (11, True)):
        pairs, d_units, d_gaps, bits = hd.encode_device(data, max_len, pad_unit=selfsync)
        assert bits > (1 << 34)
        tab = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11)) if selfsync else hd.Codetable(pairs=pairs)
        dec = hd.Decoder(d_units.numel(), selfsync=selfsync)
        out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        if selfsync:
            dec.decode_selfsync(d_units, tab, n, out=out)
        else:
            dec.decode_gap(d_units, d_gaps, tab, n, out=out)
        torch.cuda.synchronize()
        assert bool(torch.equal(out, data))
        tab.close()
        del out, dec, d_units, d_gaps
