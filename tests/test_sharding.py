"""Host logic of the multi-GPU sharding: shard plan, seam check, output placement -- on CPU, including
a world_size-2 run over torch.distributed/gloo where each rank "decodes" its shard with the CPU oracle."""
import os
import socket

import numpy as np
import pytest

import _oracle as O
import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plan_covers_stream_on_tile_boundaries(hd):
    for n_units in (1, 5, 1023, 1024, 1025, 4 * 256 * 7 + 3, 19_500_001):
        for world in (1, 2, 3, 4, 8):
            plan = hd.plan_shards(n_units, world)
            n_seg = (n_units + 3) // 4
            assert plan[0][0] == 0 and plan[-1][1] == n_seg
            for g, (b, e, h) in enumerate(plan):
                assert b % 256 == 0 and h % 8 == 0 and h <= b
                if g:
                    assert b == plan[g - 1][1]
                sb, se = hd.shard_segments(n_seg, g, world)       # C-ABI helper agrees with the Python plan
                assert (sb, se) == (b, e)


def test_plan_shard_fields_both_schemes(hd):
    """hd_mgpu_plan_shard: units / gap words a rank uploads (pure host arithmetic of the C ABI)."""
    n_units = 4 * 256 * 9 + 2                      # 10 tiles, the last one partial
    n_seg = (n_units + 3) // 4
    for world in (1, 2, 3, 8):
        covered = 0
        for r in range(world):
            g = hd.plan_shard(hd.SCHEME_GAP, n_units, r, world, 3)
            s = hd.plan_shard(hd.SCHEME_SELFSYNC, n_units, r, world, 3)
            assert (g.seg_begin, g.seg_end) == (s.seg_begin, s.seg_end) == hd.shard_segments(n_seg, r, world)
            assert g.halo_units == 0 and g.first_unit == g.seg_begin * 4
            assert g.shard_units == min(g.seg_end * 4, n_units) - g.seg_begin * 4
            assert g.first_unit + g.readable_units <= n_units and g.readable_units >= g.shard_units
            # gap words: the one before the shard (entry offset of its first segment) up to the one holding seg_end-1
            assert g.first_gap_word == (g.seg_begin // 8 - 1 if g.seg_begin else 0)
            assert g.first_gap_word + g.n_gap_words == (g.seg_end + 7) // 8
            assert s.halo_units % 32 == 0 and s.first_unit == g.first_unit - s.halo_units
            assert (s.halo_units == 0) == (r == 0)
            assert g.out_capacity >= (g.shard_units * 32 + 2) // 3
            covered += g.shard_units
        assert covered == n_units


def test_plan_shard_more_ranks_than_tiles(hd):
    """a stream with fewer 256-segment tiles than ranks leaves ranks without segments; they take part in the
    exchange with an 'empty' triple and the seam check chains through them."""
    n_units = 4 * 256 + 8                          # two tiles
    plans = [hd.plan_shard(hd.SCHEME_SELFSYNC, n_units, r, 4, 1) for r in range(4)]
    assert sum(int(p.shard_units) for p in plans) == n_units
    assert sum(1 for p in plans if p.shard_units == 0) == 2
    E = (1 << 64) - 1
    infos = [(100, 0, 5), (0, E, E), (40, 5, 0), (0, E, E)]
    assert hd.seams_ok(infos) == []
    assert hd.seams_ok([(100, 0, 5), (0, E, E), (40, 6, 0), (0, E, E)]) == [2]
    assert hd.place_outputs([100, 0, 40, 0], 130) == [(0, 100), (100, 0), (100, 30), (130, 0)]


def test_seam_check_ignores_the_unconverged_flag(hd):
    """bit 32 of the entry word = "decoded before the synchronisation had converged" (hd_mgpu_finish repeats those)."""
    assert hd.seams_ok([(5, 0, 3), (6, 3 | (1 << 32), 7), (2, 7, 0)]) == []


def test_place_outputs_clamps_last_shard(hd):
    assert hd.place_outputs([10, 20, 35], 60) == [(0, 10), (10, 20), (30, 30)]
    assert hd.place_outputs([10, 0, 5], 12) == [(0, 10), (10, 0), (10, 2)]


def test_seams_ok(hd):
    assert hd.seams_ok([(5, 0, 3), (6, 3, 7), (2, 7, 0)]) == []
    assert hd.seams_ok([(5, 0, 3), (6, 2, 7), (2, 6, 0)]) == [1, 2]


def _oracle_shard(units, lut, n_symbols, plan, g, entry=None):
    """What one rank's CUDA shard decode reports, computed with the oracle: (bytes, (count, entry, exit))."""
    start, counts = O.sync_points(units, lut, n_symbols, 128)     # true parse of the whole stream
    b, e, h = plan[g]
    true_entry = int(start[b])
    if entry is None:
        entry = true_entry
    assert entry == true_entry, "test only models repairs that land on the true parse"
    full = O.decode_flat(units, lut, int(counts.sum()))
    o = int(counts[:b].sum())
    c = int(counts[b:e].sum())
    exit_off = int(start[e]) if e < len(start) else 0
    return full[o:o + c], (c, entry, exit_off)


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    import hd_b200 as hd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 300_000
        data = O.generate(n, 0.5, seed=11)
        syms, lens = O.huffman_lengths(data, 11)
        pairs = cases.pairs_of(syms, lens)
        units, gw, bits, lut = cases.oracle_stream(data, pairs, pad_unit=True)
        plan = hd.plan_shards(len(units), world)
        out, info = _oracle_shard(units, lut, n, plan, rank)
        # rank 1 pretends its halo was too short: wrong entry offset, wrong bytes
        state = {"out": out, "info": info, "redone": 0}
        if rank == 1:
            state["info"] = (info[0] + 1, (info[1] + 1) % 16, info[2])

        def gather():
            t = torch.tensor(state["info"], dtype=torch.int64)
            lst = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(lst, t)
            return np.stack([x.numpy() for x in lst])

        def redecode(first_bit):
            o, i = _oracle_shard(units, lut, n, plan, rank, entry=first_bit)
            state["out"], state["info"], state["redone"] = o, i, state["redone"] + 1

        placed = hd.resolve_seams(rank, world, gather(), gather, redecode, n)
        off, cnt = placed[rank]
        assert state["redone"] == (1 if rank == 1 else 0)
        assert np.array_equal(state["out"][:cnt], data[off:off + cnt])
        assert sum(c for _, c in placed) == n
        with open(os.path.join(tmp, "ok%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_seam_protocol_world2_gloo(hd, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
