"""PCIe floor of the multi-GPU end-to-end path: every rank copies its shard's bytes (pinned host <-> its own GPU) at the same
time as all the other ranks -- H2D alone, D2H alone, both directions at once.  Run under torchrun (one process per GPU):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_mgpu.py [MiB_in MiB_out]
The aggregate over ranks is what the sharded e2e number of bench.py can reach at best on this box."""
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in = (int(sys.argv[1]) if len(sys.argv) > 1 else 761) << 20      # 1 GiB at redundancy 0.5: 761 MiB of stream
n_out = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024) << 20
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(2)
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if rank == 0:
    print("ranks %d, per rank: %d MiB up, %d MiB down; host cores %d" % (world, n_in >> 20, n_out >> 20, os.cpu_count()))
    for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "lscpu | grep -i -E 'numa|model name|socket'"]):
        try:
            print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout)
        except Exception as e:   # noqa: BLE001
            print("(%s: %s)" % (" ".join(cmd), e))
run(True, True)
for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both directions", True, True)):
    t = run(a, b)
    if rank == 0:
        up = world * n_in / t / 1e9 if a else 0.0
        dn = world * n_out / t / 1e9 if b else 0.0
        print("%-16s %8.3f ms   aggregate H2D %6.1f GB/s  D2H %6.1f GB/s   (decoded bytes / time = %.1f GB/s)"
              % (name, t * 1e3, up, dn, world * n_out / t / 1e9 if b else 0.0), flush=True)
if world > 1:
    dist.destroy_process_group()
