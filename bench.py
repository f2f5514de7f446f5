#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Huffman decoder (contract: one JSON line on rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's own CPU decoder on host cores

Metric (BASELINE.json): decoded GB/s, device-timed.  A "step" = one complete decode of the workload
(self-synchronisation scheme, `./bin/demo 0 100` shape: 100 MiB of generate.cpp data at redundancy 0.5 per GPU,
llhuff <=11-bit canonical code, 128-bit subsequences).  At N>1 ONE stream of N x 100 MiB is sharded on
subsequence boundaries (weak scaling: per-GPU work fixed); each rank decodes its shard independently and only the
per-rank symbol counts / seam offsets are exchanged inside the timed region (peer-memory mailbox over NVLink,
hd_mailbox_*; NCCL all-gather where CUDA IPC is not available).  --mib / --redundancy select the other BASELINE sizes
(e.g. --gpus 8 --mib 1024: the 8 GB stream of configs[3]).

  value     n_symbols(all ranks) / device time (CUDA events on the decode stream, max over ranks), inputs in HBM
  e2e       same metric through the host-buffer C-ABI call (hd_host_decode_selfsync): pinned host -> H2D -> decode
            -> D2H inside the timed region
  roofline  dominant kernel (K4 decode-and-write): algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline  the reference's sequential.cpp decode (oracle/_ref build) on a bounded sample, 1 host core
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def workload(args):
    """configs[1] of BASELINE.json by default; --mib / --redundancy name the other sizes (e.g. --gpus 8 --mib 1024 =
    configs[3], the 8 GB stream sharded over the box)."""
    return "gpuhd self-sync decode, %d MiB generate.cpp r=%g per GPU (./bin/demo 0 %d shape)" % (args.mib, args.redundancy, args.mib)
METRIC = "decoded GB/s (device-timed)"
FLUSH_BYTES = 256 << 20   # > 126 MB L2


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ----------------------------------------------------------------------------------------------------------
# clocks: sampled through NVML during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.0005)

    def sample(self):
        """One synchronous sample (called right after a step has been enqueued, i.e. while the GPU executes it)."""
        if not self.nv:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------------------
def _sample_data(mib, r, seed=1):
    import hd_b200 as hd
    return hd.generate(mib << 20, r, seed=seed)


def run_ref_program(binary, data, pattern):
    """Run an UNMODIFIED reference CPU program (built by oracle/Makefile into oracle/_ref) on `data`; they read a
    hard-coded file name from the cwd (sequential.cpp:240, parallel_cpu_prescan.cpp:596)."""
    d = tempfile.mkdtemp(prefix="hd_ref_")
    try:
        data.tofile(os.path.join(d, "data.bin"))
        os.link(os.path.join(d, "data.bin"), os.path.join(d, "data100_100.bin"))
        out = subprocess.run([binary], cwd=d, capture_output=True, text=True, timeout=1200).stdout
    finally:
        shutil.rmtree(d, ignore_errors=True)
    m = re.search(pattern, out)
    ok = "PASS" in out
    if not m:
        raise RuntimeError("could not parse reference output:\n" + out[-2000:])
    return float(m.group(1)) * 1e-6, ok


def cpu_baseline_sequential(r, sample_mib):
    """sequential.cpp HuffmanSequential::decode (sequential.cpp:57-97) on one host core."""
    ref = os.path.join(ROOT, "oracle", "_ref", "sequential")
    data = _sample_data(sample_mib, r)
    if os.path.exists(ref):
        secs, ok = run_ref_program(ref, data, r"Decompression time:\s+(\d+) mcs")
        return {"value": round(len(data) / secs / 1e9, 6), "unit": "GB/s", "cores": 1, "kind": "reference",
                "sample": "%d MiB of the same generate.cpp r=%.2f data through oracle/_ref/sequential "
                          "(sequential.cpp decode, Verification %s)" % (sample_mib, r, "PASS" if ok else "FAIL")}
    # oracle port: flat-LUT sequential decode
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    import hd_b200 as hd
    pairs = hd.build_code(data, 11)
    lut = hd.flat_lut(pairs, 11)
    units, _ = hd.encode(data, pairs, pad_unit=True)
    t0 = time.time()
    got = O.decode_flat(units, lut, len(data))
    secs = time.time() - t0
    assert np.array_equal(got, data)
    return {"value": round(len(data) / secs / 1e9, 6), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": "%d MiB, oracle flat-LUT sequential decode (oracle/hd_oracle.c)" % sample_mib}


def reference_arm(args):
    """The reference's best CPU decoder for this path on all host cores: parallel_cpu_prescan.cpp
    (compute_safe_points + decode_parallel_safe, OpenMP omp_get_max_threads())."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    binary = os.path.join(ROOT, "oracle", "_ref", "parallel_decomp_cpu")
    sample_mib = args.ref_sample_mib
    data = _sample_data(sample_mib, args.redundancy)
    times = []
    kind = "reference"
    if os.path.exists(binary):
        what = "oracle/_ref/parallel_decomp_cpu (parallel_cpu_prescan.cpp, OpenMP %d threads)" % cores
        for i in range(args.warmup + args.steps):
            secs, ok = run_ref_program(binary, data, r"Decompression time \(Parallel Decode\):\s+(\d+) us")
            if not ok:
                raise RuntimeError("reference program failed its own verification")
            if i >= args.warmup:
                times.append(secs)
    else:
        kind = "port"
        what = "oracle flat-LUT sequential decode (oracle/hd_oracle.c), reference binaries absent"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle as O
        import hd_b200 as hd
        pairs = hd.build_code(data, 11)
        lut = hd.flat_lut(pairs, 11)
        units, _ = hd.encode(data, pairs, pad_unit=True)
        cores = 1
        for i in range(args.warmup + args.steps):
            t0 = time.time()
            O.decode_flat(units, lut, len(data))
            if i >= args.warmup:
                times.append(time.time() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = len(data) / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload(args), "scheme": "cpu", "redundancy": args.redundancy},
        "cpu_baseline": {"value": round(value, 6), "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": "each step decodes a %d MiB sample of the workload with %s; host CPU only"
                                   % (sample_mib, what)},
        "e2e": {"value": round(value, 6), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def build_inputs(args, rank, world):
    """Rank 0 generates + encodes the whole stream (host threads), every rank maps its shard from /dev/shm."""
    import hd_b200 as hd
    n_total = (args.mib << 20) * world
    tag = "hd_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid() if world > 1 else os.getpid())
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    paths = {k: os.path.join(shm, "%s_%s.npy" % (tag, k)) for k in ("data", "units", "lut", "pairs")}
    if world == 1:
        data = hd.generate(n_total, args.redundancy, seed=1)
        pairs = hd.build_code(data, 11)
        lut = hd.flat_lut(pairs, 11)
        units, _ = hd.encode(data, pairs, pad_unit=True)
        return data, units, lut, pairs, None
    import torch.distributed as dist
    if rank == 0:
        data = hd.generate(n_total, args.redundancy, seed=1)
        pairs = hd.build_code(data, 11)
        lut = hd.flat_lut(pairs, 11)
        units, _ = hd.encode(data, pairs, pad_unit=True)
        for k, v in (("data", data), ("units", units), ("lut", lut), ("pairs", pairs)):
            np.save(paths[k], v)
    dist.barrier()
    data = np.load(paths["data"], mmap_mode="r")
    units = np.load(paths["units"], mmap_mode="r")
    lut = np.load(paths["lut"])
    pairs = np.load(paths["pairs"])
    return data, units, lut, pairs, paths


def ours(args):
    import torch
    import hd_b200 as hd
    hd.lib()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the decoder has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    data, units, lut, pairs, paths = build_inputs(args, rank, world)
    n_total = len(data)
    table = hd.Codetable(flat_lut=lut)

    if world == 1:
        d_units = hd.to_device_units(units)
        n_sym = n_total
        dec = hd.Decoder(d_units.numel(), selfsync=True)
        out = torch.empty(n_sym, dtype=torch.uint8, device=dev)

        def step():      # everything on the stream, no host synchronisation (hd_decode_selfsync_enqueue)
            dec.decode_selfsync(d_units, table, n_sym, out=out, defer=True)

        def finish():    # stream sync + check of the 4-byte convergence flag (hd_decode_selfsync_finish)
            dec.finish()
        shard = None
    else:
        from hd_b200 import ShardedSelfSyncDecoder
        shard = ShardedSelfSyncDecoder(units, table, n_total, rank, world, dist=dist)
        step = shard.step      # enqueue the shard decode + the exchange of the counts, no host synchronisation

        def finish():          # the sharded steps are checked together after the timed loop (shard.wait): a host
            pass               # sync per step would put the launch skew between the 8 processes into every step

    flush = torch.zeros(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.add_(1)
        if shard is not None:
            shard.align()
        step()
        finish()
    if shard is not None:
        shard.wait()
    barrier()

    # -------- parity of what we are about to time (outside the timed region)
    if world == 1:
        ok = bool(torch.equal(out, torch.from_numpy(np.ascontiguousarray(data)).to(dev)))
    else:
        ok = shard.verify(data)
    if dist is not None:
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    if not ok:
        raise SystemExit("bench.py: decoded output differs from the original data -- not timing a wrong result")

    # -------- timed region: exactly K steps, device-timed per step (L2 flushed between steps, outside the events)
    clocks = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = hd.launch_count()
    barrier()
    clocks.start()
    for i in range(args.steps):
        flush.add_(1)
        if shard is not None:
            shard.align()   # the flush is not part of the workload: ranks whose flush took longer are waited for HERE,
                            # outside the events, so that the timed exchange only sees the skew of the decode itself
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        if world == 1:
            clocks.sample()   # NVML query while the GPU is executing this step (does not touch the stream); at N > 1
                              # the steps are enqueued back to back and only the sampler thread watches the clocks
        finish()   # host-side check of the step's convergence flag, after the device-timed interval
    if shard is not None:
        shard.wait()          # stream sync + the steps' convergence flags (collective if some rank decoded again)
    barrier()
    launches = hd.launch_count() - launches0
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    if dist is not None:
        t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
    ms_per_step = t_ms / args.steps
    value = n_total / (ms_per_step * 1e-3) / 1e9

    # -------- per-kernel device time (separate pass with the library's event hooks) -> roofline of K4
    kern = {k: 0.0 for k in hd.PROFILE_SLOTS}
    for i in range(args.steps):
        flush.add_(1)
        hd.profile_enable(True)
        step()
        finish()
        for k, v in hd.profile_read().items():
            kern[k] += v / args.steps
    hd.profile_enable(False)
    if shard is not None:
        shard.wait()
    barrier()
    clk = clocks.stop()   # sampled over the timed region and the per-kernel pass of the same workload
    if world == 1:
        n_units_local, n_sym_local = len(units), n_total
    else:
        n_units_local, n_sym_local = shard.n_units_local, shard.count_local
    n_seg_local = (n_units_local + 3) // 4
    algo_bytes = 4 * n_units_local + n_sym_local + 4 * ((n_seg_local + 7) // 8)
    peak, peak_src = measured_peak()
    k4_ms = kern["decode_write"]
    achieved = algo_bytes / (k4_ms * 1e-3) / 1e9 if k4_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k4_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch_100MiB")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "hd_decode_write_kernel (K4)", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": round(k4_ms, 4), "peak_source": peak_src,
                "step_roofline_frac": round((4 * n_units_local + n_sym_local) / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                "kernels_ms": {k: round(v, 4) for k, v in kern.items() if k != "reserved"}}

    # -------- e2e: host buffers through the C ABI, H2D + decode + D2H timed (device events + host wall clock)
    e2e = None
    if world == 1:
        h_units = torch.from_numpy(np.ascontiguousarray(units).view(np.int32)).pin_memory()
        h_out = torch.empty(n_total, dtype=torch.uint8).pin_memory()
        hdec = hd.HostDecoder(len(units), n_total, device=local)
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            hdec.decode_selfsync(h_units, lut, h_out, n_total)
        tot = 0.0
        for _ in range(e_steps):
            t0 = time.perf_counter()
            hdec.decode_selfsync(h_units, lut, h_out, n_total)
            tot += time.perf_counter() - t0
        ok2 = bool(np.array_equal(h_out.numpy(), data))
        hdec.close()
        if not ok2:
            raise SystemExit("bench.py: host-buffer decode differs from the original data")
        e2e = {"value": round(n_total / (tot / e_steps) / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": int(4 * len(units) + lut.nbytes), "d2h_bytes_per_step": int(n_total),
               "steps": e_steps, "api": "hd_host_decode_selfsync (pinned host in/out, wall clock)"}
    else:
        e2e = shard.e2e(data, steps=max(3, min(args.steps, 10)))

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload(args), "scheme": "selfsync", "mib_per_gpu": args.mib, "redundancy": args.redundancy,
                   "subsequence_bits": 128, "max_codeword_length": 11, "n_symbols_total": int(n_total),
                   "n_units_total": int(len(units)), "l2": "flushed between steps (256 MiB write, outside the events)",
                   "parallelism": "shard%d" % world,
                   "exchange": ("none" if world == 1 else
                                "peer-memory mailbox over NVLink (hd_mailbox_*)" if shard.mailbox is not None else
                                "ncclAllGather of 24 B")},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity": "byte-exact",
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sequential(args.redundancy, args.cpu_sample_mib)
        print(json.dumps(line), flush=True)
    if paths and rank == 0:
        barrier()
        for p in paths.values():
            try:
                os.remove(p)
            except OSError:
                pass
    elif dist is not None:
        barrier()
    if shard is not None:
        shard.close()
    table.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=100, help="original MiB per GPU (demo.cc: MB = 2^20 B)")
    ap.add_argument("--redundancy", type=float, default=0.5)
    ap.add_argument("--cpu-sample-mib", type=int, default=32)
    ap.add_argument("--ref-sample-mib", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
