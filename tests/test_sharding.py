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
