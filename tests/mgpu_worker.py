"""torchrun worker of tests/test_multigpu.py: one process per GPU, real CUDA shards through hd_mgpu_* with both
exchanges (peer-memory mailbox; NCCL all-gather through the callback) and both schemes.  Test infrastructure."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import hd_b200 as hd
    import _oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_003
    data = O.generate(n, 0.5, seed=77)                       # every rank makes the same seeded input
    # self-sync scheme (gpuhd layout, 11-bit code)
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, _ = hd.encode(data, pairs, pad_unit=True)
    tab = hd.Codetable(flat_lut=lut)
    for use_mailbox in (True, False):
        d = hd.ShardedSelfSyncDecoder(units, tab, n, rank, world, dist=dist, use_mailbox=use_mailbox)
        assert d.verify(data), "self-sync shard mismatch (mailbox=%s)" % use_mailbox
        for _ in range(3):                                    # steps enqueued back to back, one collective finish
            d.step()
        placed = d.finish()
        assert sum(c for _, c in placed) == n and d.repairs() == 0
        e2e = d.e2e(data, steps=2)                            # pinned host shard in/out through hd_mgpu_host_step
        assert e2e["d2h_bytes_per_step"] == n
        whole = d.gather_output(0)
        if rank == 0:
            assert bool(torch.equal(whole, torch.from_numpy(data).cuda())), "gathered output differs"
            print("selfsync ok: exchange=%s" % ("mailbox" if d.mailbox is not None else "nccl-callback"), flush=True)
        d.close()
    tab.close()
    # gap-array scheme (.huff layout, 16-bit code)
    pairs16 = hd.build_code(data, 16)
    units16, gaps, _ = hd.encode(data, pairs16, want_gaps=True)
    tab16 = hd.Codetable(pairs=pairs16)
    g = hd.ShardedGapDecoder(units16, gaps, tab16, n, rank, world, dist=dist)
    assert g.verify(data), "gap shard mismatch"
    e2e = g.e2e(data, steps=2)
    if rank == 0:
        print("gap ok: e2e %.1f GB/s" % e2e["value"], flush=True)
    g.close()
    tab16.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_WORKER_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
