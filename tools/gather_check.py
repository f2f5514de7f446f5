"""torchrun check: sharded decode on N GPUs, output gathered on rank 0 over NVLink, compared with the original."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd
from hd_b200 import ShardedSelfSyncDecoder

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
n = (256 << 20) + 12345
data = hd.generate(n, 0.5, seed=3)                 # every rank builds the same stream (seeded)
pairs = hd.build_code(data, 11)
units, bits = hd.encode(data, pairs, pad_unit=True)
table = hd.Codetable(flat_lut=hd.flat_lut(pairs, 11))
sh = ShardedSelfSyncDecoder(units, table, n, rank, world, dist=dist)
sh.step(); sh.finish()
whole = sh.gather_output(0)                        # first use sets up the NCCL point-to-point connections
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
whole = sh.gather_output(0)
torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter() - t0
if rank == 0:
    ok = bool(torch.equal(whole, torch.from_numpy(data).cuda()))
    print("gathered %d bytes from %d ranks on rank 0 in %.2f ms (%.0f GB/s over NVLink): identical=%s" % (n, world, t * 1e3, n * (world - 1) / world / t / 1e9, ok))
    assert ok
sh.close()
dist.destroy_process_group()
