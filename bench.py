#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Huffman decoder (contract: one JSON line on rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's own CPU decoder on host cores

Metric (BASELINE.json): decoded GB/s, device-timed.  A "step" = one complete decode of the workload.

HEADLINE (`value`): self-synchronisation scheme (gpuhd, `./bin/demo <gpu> <MB>` shape), 1 GiB of generate.cpp data at
redundancy 0.5 PER GPU, llhuff <= 11-bit canonical code, 128-bit subsequences -- the largest single-GPU configuration of
BASELINE.json; with --gpus 8 this is configs[3], ONE 8 GiB stream sharded over the box on subsequence boundaries (weak
scaling: per-GPU work fixed).  Every rank decodes its shard independently; only {count, entry, exit} per rank are
exchanged inside the timed region (hd_mgpu_*: peer-memory mailbox over NVLink, NCCL all-gather where CUDA IPC is not
available).

Inputs come from the REFERENCE's own producers when oracle/_ref (built by oracle/Makefile from /root/reference) is there:
generate.cpp (oracle/_ref/generate, several processes), the llhuff encoder (oracle/_ref/libref_llhuff.so) and the
gap-array GPU encoder (oracle/_ref/gap_encoder); `config.input` says what was used.  They only PRODUCE inputs; nothing
under oracle/ is on a timed path of our arm.

  value        n_symbols(all ranks) / device time (CUDA events on the decode stream, max over ranks), inputs in HBM
  e2e          same metric through the host-buffer C-ABI call: pinned host -> H2D -> decode -> D2H in the timed region
  roofline     dominant kernel (K4 decode-and-write): algorithmic bytes / its CUDA-event duration vs measured HBM peak;
               step_roofline_frac = the whole decode (all kernels) against the same peak
  records      the other BASELINE.json configurations, each with its own roofline: configs[1] self-sync 100 MiB,
               configs[2] gap-array 1 GiB (.huff), configs[4] redundancy sweep at 1 GiB (both schemes); at N > 1 a
               STRONG-scaling point (one 1 GiB stream over N GPUs) and the sharded gap-array decode
  cpu_baseline the reference's sequential.cpp decode on a bounded sample (1 core); cpu_baselines adds parallel_cpu.cpp
               and parallel_cpu_prescan.cpp (OpenMP)
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
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)

METRIC = "decoded GB/s (device-timed)"
FLUSH_BYTES = 256 << 20   # > 126 MB L2
SWEEP = (0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 1.0)


def workload(mib, r, world=1):
    return ("gpuhd self-sync decode, %d MiB generate.cpp r=%g per GPU (./bin/demo <gpu> %d shape)%s"
            % (mib, r, mib, "" if world == 1 else "; one %d MiB stream sharded over %d GPUs" % (mib * world, world)))


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


def scratch_dir():
    return "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()


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

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(0.0005)

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
# input producers (the reference's own, when oracle/_ref is there)
# ----------------------------------------------------------------------------------------------------------
def numpy_generate(n, r, seed=1):
    """The distribution of generate.cpp:33-47 with numpy (used only when neither oracle/_ref/generate nor the product
    library may be touched: the reference arm without a reference build)."""
    rng = np.random.default_rng(seed)
    hot = rng.random(n) < r
    out = rng.integers(0, 256, n, dtype=np.uint8)
    out[hot] = (ord("A") + rng.integers(0, 4, int(hot.sum()))).astype(np.uint8)
    return out


def ref_generate(n, r, procs=0):
    """n bytes from the UNMODIFIED reference generator (generate.cpp, built as oracle/_ref/generate): `procs` processes,
    each in its own directory because the program always writes ./data.bin (generate.cpp:24) and seeds itself from
    random_device (generate.cpp:32).  None if the binary is not there."""
    exe = os.path.join(REF, "generate")
    if not os.path.exists(exe) or n == 0:
        return None
    procs = procs or max(1, min(32, os.cpu_count() or 1, (n + (1 << 22) - 1) >> 22))
    piece = (n + procs - 1) // procs
    base = tempfile.mkdtemp(prefix="hd_gen_", dir=scratch_dir())
    try:
        jobs = []
        for i in range(procs):
            size = min(piece, n - i * piece)
            if size <= 0:
                break
            d = os.path.join(base, "p%d" % i)
            os.mkdir(d)
            jobs.append((d, size, subprocess.Popen([exe, str(size), repr(float(r))], cwd=d, stdout=subprocess.DEVNULL)))
        data = np.empty(n, dtype=np.uint8)
        off = 0
        for d, size, pr in jobs:
            if pr.wait() != 0:
                return None
            data[off:off + size] = np.fromfile(os.path.join(d, "data.bin"), dtype=np.uint8, count=size)
            os.remove(os.path.join(d, "data.bin"))
            off += size
        return data
    finally:
        shutil.rmtree(base, ignore_errors=True)


def make_data(n, r, allow_product=True):
    data = ref_generate(n, r)
    if data is not None:
        return data, "reference generate.cpp (oracle/_ref/generate, unmodified, parallel processes)"
    if allow_product:
        import hd_b200 as hd
        return hd.generate(n, r, seed=1), "product restatement of generate.cpp (hd_generate, seeded; oracle/_ref absent)"
    return numpy_generate(n, r), "numpy restatement of generate.cpp:33-47 (oracle/_ref absent)"


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    return O


def selfsync_stream(data, full_ref_limit=1 << 30, head=256 << 20):
    """gpuhd stream of `data`: (units incl. pad unit, flat LUT 2048x2, description).  Up to 1 GiB the UNMODIFIED
    reference encoder does everything (get_symbol_lengths, get_encoder_table, get_decoder_table, encode_memory,
    CUHDInputBuffer pad).  Beyond that (it is single-threaded: minutes for 8 GiB) the reference builds the code and LUT
    from the first 256 MiB, the product's multi-threaded packer packs the stream with that code, and the packed prefix is
    checked bit for bit against the reference's encode_memory output."""
    import hd_b200 as hd
    O = _oracle()
    if not O.have_ref():
        pairs = hd.build_code(data, 11)
        units, _ = hd.encode(data, pairs, pad_unit=True)
        return units, hd.flat_lut(pairs, 11), "product code builder + packer (hd_build_code / hd_encode; oracle/_ref absent)"
    if len(data) <= full_ref_limit:
        ref = O.ref_llhuff_encode(data)
        return ref["units"], ref["lut"], "reference llhuff encoder, unmodified (oracle/_ref/libref_llhuff.so: tables + encode_memory + pad unit)"
    h = min(head, len(data))
    ref = O.ref_llhuff_encode(data[:h])
    lens, codes = ref["lens"].astype(np.int64), ref["codes"].astype(np.int64)
    present = np.nonzero(lens)[0]
    order = sorted(present, key=lambda s: (lens[s], codes[s]))
    pairs = np.array([[s, lens[s]] for s in order], dtype=np.uint8)
    lut = hd.flat_lut(pairs, 11)
    units = None
    if not len(np.setdiff1d(np.unique(data[h:h + (64 << 20)]), present)) and np.array_equal(lut, ref["lut"]):
        units, _ = hd.encode(data, pairs, pad_unit=True)
        k = int(lens[data[:h]].sum()) // 32
        if not np.array_equal(units[:k], ref["units"][:k]):
            units = None
    if units is None:   # (never seen: the head's code does not cover the stream, or is not canonical in list order)
        pairs = hd.build_code(data, 11)
        units, _ = hd.encode(data, pairs, pad_unit=True)
        return units, hd.flat_lut(pairs, 11), "product code builder + packer (the reference's head code did not carry over)"
    return units, lut, ("reference llhuff code + LUT (unmodified, from the first %d MiB); product multi-threaded packer, "
                        "bit-identical to the reference's encode_memory on that prefix (%d units checked)" % (h >> 20, k))


def gap_stream(data, try_ref=True):
    """.huff container of `data`: (pairs, units, gap words, description).  The UNMODIFIED reference GPU encoder
    (Huffman_coding_Gap_arrays/encoder, oracle/_ref/gap_encoder) writes it when it runs here (it needs this GPU and
    < 4 GiB of input); otherwise the product's host producer (same container, bit-exact layout)."""
    import hd_b200 as hd
    exe = os.path.join(REF, "gap_encoder")
    if try_ref and os.path.exists(exe) and len(data) < (1 << 32):
        d = tempfile.mkdtemp(prefix="hd_huff_", dir=scratch_dir())
        try:
            src, dst = os.path.join(d, "in.bin"), os.path.join(d, "out.huff")
            data.tofile(src)
            run = subprocess.run([exe, src, dst], capture_output=True, text=True, timeout=600)
            if run.returncode == 0 and os.path.exists(dst):
                hf = hd.HuffFile(np.fromfile(dst, dtype=np.uint8))
                if hf.original_size == len(data):
                    return (hf.pairs, hf.units, hf.gap_words,
                            "reference gap-array GPU encoder, unmodified (oracle/_ref/gap_encoder) -> .huff")
        except Exception:
            pass
        finally:
            shutil.rmtree(d, ignore_errors=True)
    pairs = hd.build_code(data, 16)
    units, gaps, _ = hd.encode(data, pairs, want_gaps=True)
    return pairs, units, gaps, "product .huff producer (hd_build_code 16-bit limit + hd_encode with gap array)"


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baselines (the places that TIME something under oracle/)
# ----------------------------------------------------------------------------------------------------------
def run_ref_program(binary, data, pattern, threads=None):
    """Run an UNMODIFIED reference CPU program (built by oracle/Makefile into oracle/_ref) on `data`; they read a
    hard-coded file name from the cwd (sequential.cpp:240, parallel_cpu.cpp:322, parallel_cpu_prescan.cpp:596)."""
    d = tempfile.mkdtemp(prefix="hd_ref_", dir=scratch_dir())
    try:
        data.tofile(os.path.join(d, "data.bin"))
        os.link(os.path.join(d, "data.bin"), os.path.join(d, "data100_100.bin"))
        env = dict(os.environ)
        if threads:
            env["OMP_NUM_THREADS"] = str(threads)
        out = subprocess.run([binary], cwd=d, capture_output=True, text=True, timeout=1800, env=env).stdout
    finally:
        shutil.rmtree(d, ignore_errors=True)
    m = re.search(pattern, out)
    if not m:
        raise RuntimeError("could not parse reference output:\n" + out[-2000:])
    return float(m.group(1)) * 1e-6, "PASS" in out


def cpu_baselines(r, sample_mib):
    """The reference's CPU decoders on the box's host cores, each on a bounded sample of the same data (the programs
    hold one char per code bit: 100 MiB needs ~0.7 GB and ~20 s of one core)."""
    res = []
    data, src = make_data(sample_mib << 20, r)
    specs = [("sequential", "sequential.cpp HuffmanSequential::decode (sequential.cpp:57-97)", r"Decompression time:\s+(\d+) mcs", 1),
             ("parallel_cpu", "parallel_cpu.cpp (OpenMP thread_count=16 hard-coded, decode is sequential; parallel_cpu.cpp:329-334)",
              r"Decompression time:\s+(\d+) mcs", 16),
             ("parallel_decomp_cpu", "parallel_cpu_prescan.cpp compute_safe_points + decode_parallel_safe (OpenMP, :423-483)",
              r"Decompression time \(Parallel Decode\):\s+(\d+) us", os.cpu_count() or 1)]
    for exe, what, pat, cores in specs:
        binary = os.path.join(REF, exe)
        if not os.path.exists(binary):
            continue
        try:
            secs, ok = run_ref_program(binary, data, pat)
        except Exception as e:   # noqa: BLE001
            res.append({"program": exe, "error": str(e)[:200]})
            continue
        res.append({"program": exe, "value": round(len(data) / secs / 1e9, 6), "unit": "GB/s", "cores": cores,
                    "kind": "reference", "sample": "%d MiB of generate.cpp r=%.2f data (%s) through oracle/_ref/%s: %s, "
                    "Verification %s" % (sample_mib, r, src.split(" (")[0], exe, what, "PASS" if ok else "FAIL")})
    if not res:
        # oracle port: flat-LUT sequential decode
        O = _oracle()
        import hd_b200 as hd
        pairs = hd.build_code(data, 11)
        lut = hd.flat_lut(pairs, 11)
        units, _ = hd.encode(data, pairs, pad_unit=True)
        t0 = time.time()
        got = O.decode_flat(units, lut, len(data))
        secs = time.time() - t0
        assert np.array_equal(got, data)
        res.append({"program": "oracle", "value": round(len(data) / secs / 1e9, 6), "unit": "GB/s", "cores": 1, "kind": "port",
                    "sample": "%d MiB, oracle flat-LUT sequential decode (oracle/hd_oracle.c)" % sample_mib})
    return res


def reference_arm(args):
    """The reference's best CPU decoder for this path on all host cores: parallel_cpu_prescan.cpp
    (compute_safe_points + decode_parallel_safe, OpenMP omp_get_max_threads()).  Does not load the product library."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    binary = os.path.join(REF, "parallel_decomp_cpu")
    sample_mib = args.ref_sample_mib
    data, src = make_data(sample_mib << 20, args.redundancy, allow_product=False)
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
        O = _oracle()
        syms, lens = O.huffman_lengths(data, 11)
        codes, clens = O.canonical_codes(syms, lens)
        lut = O.flat_lut(codes, clens, 11)
        units = O.pack(data, codes, clens)[0]
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
        "config": {"workload": workload(args.mib, args.redundancy, args.gpus), "scheme": "cpu", "redundancy": args.redundancy,
                   "sample": "each step decodes a %d MiB sample of the workload's data (CPU throughput does not depend on "
                             "the size; the program needs ~7 B of RAM per input bit)" % sample_mib,
                   "input": src},
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
class Timer:
    """Device timing of a step on torch's current stream: L2 flushed before every step (outside the events)."""

    def __init__(self, torch, dev):
        self.torch = torch
        self.flush = torch.zeros(FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def run(self, step, finish, steps, before=None, between=None):
        torch = self.torch
        stream = torch.cuda.current_stream()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            self.flush.add_(1)
            if before:
                before()
            ev[i][0].record(stream)
            step()
            ev[i][1].record(stream)
            if between:
                between()
            if finish:
                finish()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in ev]

    def kernels(self, hd, step, finish, steps):
        kern = {k: 0.0 for k in hd.PROFILE_SLOTS}
        for _ in range(steps):
            self.flush.add_(1)
            hd.profile_enable(True)
            step()
            if finish:
                finish()
            for k, v in hd.profile_read().items():
                kern[k] += v / steps
        hd.profile_enable(False)
        return kern


def ncu_traffic(mib, r, scheme):
    """dram__bytes_read.sum + dram__bytes_write.sum of K4 per launch from the committed `ncu --set full` capture of the SAME
    workload (profiles/k4_traffic.json lists them by workload); None for workloads that were not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "k4_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"].get("%s_%dMiB_r%g" % (scheme, mib, r))
    except Exception:
        return None


def roofline_of(algo_bytes, step_bytes, k4_ms, step_ms, kern, peak, peak_src, traffic=None, fixed=0):
    achieved = algo_bytes / (k4_ms * 1e-3) / 1e9 if k4_ms > 0 else 0.0
    kname = ("hd_decode_fixed_kernel<%d> (fixed-length code: every segment starts at a codeword boundary, no synchronisation, "
             "no prescan, no gap words read)" % fixed) if fixed else "hd_decode_tiles_kernel (K4 decode-and-write)"
    return {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1),
            "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": traffic,   # from the ncu capture of this workload under profiles/ (null: workload not captured)
            "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": round(k4_ms, 4), "peak_source": peak_src,
            "step_roofline_frac": round(step_bytes / (step_ms * 1e-3) / 1e9 / peak, 4),
            "kernels_ms": {k: round(v, 4) for k, v in kern.items() if k != "reserved" and v}}


def d1_equal(torch, out, ref):
    """Byte-exact, except that the reference's encode_memory may drop the tail bits of the LAST codeword (SURVEY D1,
    llhuffman_encoder.cc:213-237): then every byte but the last must match."""
    if bool(torch.equal(out, ref)):
        return True, "byte-exact"
    if out.numel() == ref.numel() and out.numel() > 1 and bool(torch.equal(out[:-1], ref[:-1])):
        return True, "byte-exact except the last symbol (reference encoder defect D1: tail of the last codeword dropped)"
    return False, "MISMATCH"


def bench_selfsync_1gpu(hd, torch, T, data, units, lut, steps, warmup, clocks=None):
    n = len(data)
    table = hd.Codetable(flat_lut=lut)
    d_units = hd.to_device_units(units)
    dec = hd.Decoder(d_units.numel(), selfsync=True)
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    step = lambda: dec.decode_selfsync(d_units, table, n, out=out, defer=True)   # noqa: E731  enqueue only
    for _ in range(max(warmup, 3)):
        T.flush.add_(1)
        step()
        dec.finish()
    ok, parity = d1_equal(torch, out, torch.from_numpy(np.ascontiguousarray(data)).cuda())
    if not ok:
        raise SystemExit("bench.py: self-sync decode differs from the original data -- not timing a wrong result")
    l0 = hd.launch_count()
    ms = T.run(step, dec.finish, steps, between=(clocks.sample if clocks else None))
    launches = hd.launch_count() - l0
    kern = T.kernels(hd, step, dec.finish, min(steps, 10))
    fixed = table.fixed_length()
    table.close()
    n_seg = (len(units) + 3) // 4
    step_bytes = 4 * len(units) + n
    algo = step_bytes if fixed else step_bytes + 4 * ((n_seg + 7) // 8)      # K4 reads the gap words K1a left
    return {"ms": sum(ms) / len(ms), "ms_median": float(np.median(ms)), "launches": launches, "kern": kern, "algo": algo,
            "step_bytes": step_bytes, "parity": parity, "fixed": fixed}


def bench_gap_1gpu(hd, torch, T, data, pairs, units, gaps, steps, warmup):
    n = len(data)
    table = hd.Codetable(pairs=pairs)
    d_units = hd.to_device_units(units)
    d_gaps = torch.from_numpy(np.ascontiguousarray(gaps).view(np.int32)).cuda()
    dec = hd.Decoder(d_units.numel(), selfsync=False)
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    step = lambda: dec.decode_gap(d_units, d_gaps, table, n, out=out)   # noqa: E731
    for _ in range(max(warmup, 3)):
        T.flush.add_(1)
        step()
    torch.cuda.synchronize()
    if not bool(torch.equal(out, torch.from_numpy(np.ascontiguousarray(data)).cuda())):
        raise SystemExit("bench.py: gap-array decode differs from the original data -- not timing a wrong result")
    ms = T.run(step, None, steps)
    kern = T.kernels(hd, step, None, min(steps, 10))
    fixed = table.fixed_length()
    table.close()
    algo = 4 * len(units) + n + (0 if fixed else 4 * len(gaps))   # (the fixed-length kernel does not read the gap array)
    return {"ms": sum(ms) / len(ms), "ms_median": float(np.median(ms)), "kern": kern, "algo": algo, "step_bytes": algo,
            "parity": "byte-exact", "fixed": fixed}


def record(name, scheme, mib, r, res, n, peak, peak_src, input_desc, extra=None):
    ms = res["ms_median"]   # (records: median of their few steps; the headline is the mean of exactly K steps)
    rec = {"config": name, "scheme": scheme, "mib": mib, "redundancy": r, "value": round(n / (ms * 1e-3) / 1e9, 2),
           "unit": "GB/s", "ms_per_step": round(ms, 4), "ms_per_step_mean": round(res["ms"], 4),
           "roofline": roofline_of(res["algo"], res["step_bytes"], res["kern"]["decode_write"], ms, res["kern"], peak, peak_src,
                                   ncu_traffic(mib, r, scheme), res.get("fixed", 0)),
           "parity": res["parity"], "input": input_desc}
    if extra:
        rec.update(extra)
    return rec


def share_arrays(arrays, rank, world, dist, tag):
    """Rank 0 made `arrays` (dict of numpy arrays); every rank maps them from /dev/shm."""
    paths = {k: os.path.join(scratch_dir(), "%s_%s.npy" % (tag, k)) for k in arrays["_keys"]} if rank == 0 else None
    keys = arrays["_keys"] if rank == 0 else None
    box = [keys]
    dist.broadcast_object_list(box, src=0)
    keys = box[0]
    paths = {k: os.path.join(scratch_dir(), "%s_%s.npy" % (tag, k)) for k in keys}
    if rank == 0:
        for k in keys:
            np.save(paths[k], arrays[k])
    dist.barrier()
    out = {k: np.load(paths[k], mmap_mode="r") for k in keys}
    return out, paths


def ours(args):
    import torch
    import hd_b200 as hd
    hd.lib()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the decoder has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = measured_peak()
    T = Timer(torch, dev)
    r = args.redundancy
    records = []

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        # ================================================================== one GPU
        n = args.mib << 20
        data, gen_src = make_data(n, r)
        units, lut, enc_src = selfsync_stream(data)
        clocks = ClockSampler(local)
        clocks.start()
        res = bench_selfsync_1gpu(hd, torch, T, data, units, lut, args.steps, args.warmup, clocks)
        clk = clocks.stop()
        ms_per_step = res["ms"]
        value = n / (ms_per_step * 1e-3) / 1e9
        roofline = roofline_of(res["algo"], res["step_bytes"], res["kern"]["decode_write"], ms_per_step, res["kern"], peak, peak_src,
                               ncu_traffic(args.mib, r, "selfsync"), res.get("fixed", 0))
        launches, parity = res["launches"], res["parity"]
        n_total, n_units_total = n, len(units)
        input_desc = {"data": gen_src, "stream": enc_src}
        exchange = "none"

        # ---- e2e: host buffers through the C ABI, H2D + decode + D2H timed (host wall clock)
        h_units = torch.from_numpy(np.ascontiguousarray(units).view(np.int32)).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        hdec = hd.HostDecoder(len(units), n, device=local)
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            hdec.decode_selfsync(h_units, lut, h_out, n)
        tot = 0.0
        for _ in range(e_steps):
            t0 = time.perf_counter()
            hdec.decode_selfsync(h_units, lut, h_out, n)
            tot += time.perf_counter() - t0
        ok2 = np.array_equal(h_out.numpy()[:-1], data[:-1])
        hdec.close()
        if not ok2:
            raise SystemExit("bench.py: host-buffer decode differs from the original data")
        e2e = {"value": round(n / (tot / e_steps) / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": int(4 * len(units) + lut.nbytes), "d2h_bytes_per_step": int(n),
               "steps": e_steps, "api": "hd_host_decode_selfsync (pinned host in/out, wall clock)"}
        del h_units, h_out

        # ---- records: the other BASELINE.json configurations
        rsteps = max(3, min(args.steps, 10))
        if not args.no_records:
            # configs[1]: gpuhd self-sync decode, 100 MiB (./bin/demo 0 100 shape)
            d100 = data[:100 << 20] if n >= (100 << 20) else data
            u100, l100, e100 = selfsync_stream(d100)
            r100 = bench_selfsync_1gpu(hd, torch, T, d100, u100, l100, max(args.steps, 10), 3)
            records.append(record("configs[1]: gpuhd self-sync decode, 100 MiB (./bin/demo 0 100 shape)", "selfsync", len(d100) >> 20,
                                  r, r100, len(d100), peak, peak_src, {"data": gen_src, "stream": e100}))
            # configs[2]: gap-array decode of a .huff, 1 GiB
            pairs, gu, gg, gsrc = gap_stream(data)
            rg = bench_gap_1gpu(hd, torch, T, data, pairs, gu, gg, rsteps, 3)
            records.append(record("configs[2]: gap-array decode of a .huff, %d MiB" % args.mib, "gap", args.mib, r, rg, n, peak,
                                  peak_src, {"data": gen_src, "stream": gsrc}))
            del gu, gg
        if not args.no_records and not args.no_sweep:
            # configs[4]: redundancy sweep at the headline size, both schemes (reference code tables, product packer)
            for rr in SWEEP:
                if rr == r:
                    d = data
                else:
                    d, _ = make_data(n, rr)
                su, sl, ssrc = selfsync_stream(d, full_ref_limit=0, head=64 << 20)
                rs = bench_selfsync_1gpu(hd, torch, T, d, su, sl, 5, 3)
                records.append(record("configs[4]: redundancy sweep", "selfsync", args.mib, rr, rs, n, peak, peak_src,
                                      {"data": gen_src, "stream": ssrc}))
                del su
                pairs, gu, gg, gsrc = gap_stream(d, try_ref=False)
                rg = bench_gap_1gpu(hd, torch, T, d, pairs, gu, gg, 5, 3)
                records.append(record("configs[4]: redundancy sweep", "gap", args.mib, rr, rg, n, peak, peak_src,
                                      {"data": gen_src, "stream": gsrc}))
                del gu, gg, d
    else:
        # ================================================================== N GPUs: one stream sharded over the box
        from hd_b200 import ShardedSelfSyncDecoder, ShardedGapDecoder
        tag = "hd_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid())
        n_total = (args.mib << 20) * world
        arrays = None
        if rank == 0:
            data, gen_src = make_data(n_total, r)
            units, lut, enc_src = selfsync_stream(data)
            arrays = {"_keys": ["data", "units", "lut"], "data": data, "units": units, "lut": lut}
            meta = [gen_src, enc_src]
        else:
            meta = [None, None]
        shared, paths = share_arrays(arrays, rank, world, dist, tag)
        dist.broadcast_object_list(meta, src=0)
        gen_src, enc_src = meta
        data, units, lut = shared["data"], shared["units"], np.array(shared["lut"])
        table = hd.Codetable(flat_lut=lut)
        shard = ShardedSelfSyncDecoder(units, table, n_total, rank, world, dist=dist)
        exchange = ("peer-memory mailbox over NVLink (hd_mgpu_exchange_mailbox)" if shard.mailbox is not None
                    else "NCCL all-gather of 24 B (hd_mgpu_exchange_callback)")

        def run_sharded(sh, data_, steps, what):
            for _ in range(3):
                T.flush.add_(1)
                sh.align()
                sh.step()
            sh.finish()
            barrier()
            ok = sh.verify(data_) if what == "gap" else _verify_d1(sh, data_)
            t = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if not int(t.item()):
                raise SystemExit("bench.py: sharded %s decode differs from the original data" % what)
            barrier()
            l0 = hd.launch_count()
            ms = T.run(sh.step, None, steps, before=sh.align)
            run_sharded.launches = hd.launch_count() - l0      # (incl. the untimed align rounds' two tiny kernels)
            sh.finish()
            barrier()
            tt = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()) / steps

        def _verify_d1(sh, data_):
            sh.step()
            placed = sh.finish()
            got = sh.out[:sh.count_local]
            want = torch.from_numpy(np.ascontiguousarray(data_[sh.offset:sh.offset + sh.count_local])).cuda()
            if sum(c for _, c in placed) != sh.n_symbols:
                return False
            last = sh.offset + sh.count_local == sh.n_symbols
            return bool(torch.equal(got, want)) or (last and sh.count_local > 1 and bool(torch.equal(got[:-1], want[:-1])))

        clocks = ClockSampler(local)
        clocks.start()
        ms_per_step = run_sharded(shard, data, args.steps, "selfsync")
        launches = run_sharded.launches
        clk = clocks.stop()
        value = n_total / (ms_per_step * 1e-3) / 1e9
        # per-kernel time of this rank's shard
        kern = T.kernels(hd, shard.step, None, min(args.steps, 10))
        shard.finish()
        nu, ns = shard.n_units_local, shard.count_local
        algo = 4 * nu + ns + 4 * (((nu + 3) // 4 + 7) // 8)
        roofline = roofline_of(algo, 4 * nu + ns, kern["decode_write"], ms_per_step, kern, peak, peak_src,
                               ncu_traffic(args.mib, r, "selfsync"))
        parity = "byte-exact (D1-aware on the last symbol)"
        n_units_total = len(units)
        input_desc = {"data": gen_src, "stream": enc_src}
        e2e = shard.e2e(data, steps=max(3, min(args.steps, 10)))
        shard.close()
        rsteps = max(3, min(args.steps, 10))
        if not args.no_records:
            # strong scaling: ONE args.mib stream over all N GPUs (total work fixed)
            n_s = args.mib << 20
            if rank == 0:
                ds = np.ascontiguousarray(data[:n_s])
                us, ls, es = selfsync_stream(ds)
                arrays = {"_keys": ["data", "units", "lut"], "data": ds, "units": us, "lut": ls}
            sh2, p2 = share_arrays(arrays if rank == 0 else None, rank, world, dist, tag + "_s")
            t2 = hd.Codetable(flat_lut=np.array(sh2["lut"]))
            s2 = ShardedSelfSyncDecoder(sh2["units"], t2, n_s, rank, world, dist=dist)
            ms2 = run_sharded(s2, sh2["data"], rsteps, "selfsync")
            records.append({"config": "strong scaling: ONE %d MiB gpuhd stream sharded over %d GPUs" % (args.mib, world),
                            "scheme": "selfsync", "scaling": "strong", "mib_total": args.mib, "redundancy": r,
                            "value": round(n_s / (ms2 * 1e-3) / 1e9, 2), "unit": "GB/s", "ms_per_step": round(ms2, 4),
                            "parity": "byte-exact (D1-aware on the last symbol)"})
            s2.close()
            t2.close()
            barrier()
            if rank == 0:
                for p in p2.values():
                    os.remove(p)
            # gap-array scheme sharded (weak: args.mib per GPU), product 16-bit .huff producer (the reference GPU encoder
            # is limited to < 4 GiB and one GPU)
            if rank == 0:
                pairs, gu, gg, gsrc = gap_stream(data, try_ref=False)
                arrays = {"_keys": ["units", "gaps", "pairs"], "units": gu, "gaps": gg, "pairs": pairs}
                meta = [gsrc]
            else:
                meta = [None]
            sh3, p3 = share_arrays(arrays if rank == 0 else None, rank, world, dist, tag + "_g")
            dist.broadcast_object_list(meta, src=0)
            t3 = hd.Codetable(pairs=np.array(sh3["pairs"]))
            s3 = ShardedGapDecoder(sh3["units"], sh3["gaps"], t3, n_total, rank, world, dist=dist)
            ms3 = run_sharded(s3, data, rsteps, "gap")
            kern3 = T.kernels(hd, s3.step, None, min(rsteps, 5))
            s3.finish()
            nu3, ns3 = s3.n_units_local, s3.count_local
            algo3 = 4 * nu3 + ns3 + 4 * (((nu3 + 3) // 4 + 7) // 8)
            e2e3 = s3.e2e(data, steps=3)
            records.append({"config": "gap-array decode sharded over %d GPUs, %d MiB per GPU (segment entry rule decoder.cu:505-507)"
                                      % (world, args.mib), "scheme": "gap", "scaling": "weak", "mib_per_gpu": args.mib,
                            "redundancy": r, "value": round(n_total / (ms3 * 1e-3) / 1e9, 2), "unit": "GB/s",
                            "ms_per_step": round(ms3, 4), "e2e": e2e3,
                            "roofline": roofline_of(algo3, algo3, kern3["decode_write"], ms3, kern3, peak, peak_src),
                            "parity": "byte-exact", "input": {"data": gen_src, "stream": meta[0]}})
            s3.close()
            t3.close()
            barrier()
            if rank == 0:
                for p in p3.values():
                    os.remove(p)
        table.close()
        barrier()
        if rank == 0:
            for p in paths.values():
                try:
                    os.remove(p)
                except OSError:
                    pass

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload(args.mib, r, world), "scheme": "selfsync", "mib_per_gpu": args.mib, "redundancy": r,
                   "subsequence_bits": 128, "max_codeword_length": 11, "n_symbols_total": int(n_total),
                   "n_units_total": int(n_units_total), "input": input_desc,
                   "l2": "flushed between steps (256 MiB write, outside the events)", "parallelism": "shard%d" % world,
                   "exchange": exchange},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity": parity,
        "records": records,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baselines(r, args.cpu_sample_mib)
            seq = [c for c in cb if c.get("program") in ("sequential", "oracle") and "value" in c]
            if seq:
                line["cpu_baseline"] = {k: v for k, v in seq[0].items() if k != "program"}
            line["cpu_baselines"] = cb
        print(json.dumps(line), flush=True)
    if dist is not None:
        barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="original MiB per GPU (demo.cc: MB = 2^20 B)")
    ap.add_argument("--redundancy", type=float, default=0.5)
    ap.add_argument("--cpu-sample-mib", type=int, default=16)
    ap.add_argument("--ref-sample-mib", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-records", action="store_true", help="headline only")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] redundancy sweep records")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
