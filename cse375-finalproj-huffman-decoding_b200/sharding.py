"""Multi-GPU sharding of ONE Huffman stream: thin Python callers of the C ABI's hd_mgpu_* layer (one process per GPU,
torch.distributed only for the set-up plumbing).

Replaces gpuhd/multigpu_demo.cc:178-227,255-313, which slices the compressed units AND the symbols into equal 1/n
parts (wrong by construction: the reference README says "MULTI-GPU BUT TESTS FAIL", README.md:37).  The shard plan,
the decode + exchange step, the seam protocol and the output placement all live in the library
(csrc/hd_mgpu.cu): the stream is cut on 128-bit subsequence boundaries (whole 256-segment tiles), every rank decodes
its shard independently from local HBM, and the only exchange is three integers per rank

    {symbols decoded, entry bit offset assumed, bit offset at which the parse leaves the shard}

entry[g] must equal exit[g-1] (it does once the left halo is longer than the self-synchronisation distance; a
mismatching rank decodes again from the exact offset), and the exclusive scan of the counts places the outputs.
Both schemes: ShardedSelfSyncDecoder (gpuhd streams) and ShardedGapDecoder (.huff gap arrays, no halo, no repair).
"""
import ctypes as C

import numpy as np

HALO_SEGMENTS = 4096          # left halo: 4096 x 128 bit = 64 KiB of stream (>> observed max sync distance ~5 Kbit)
SEGMENT_UNITS = 4
SCHEME_SELFSYNC, SCHEME_GAP = 0, 1
EMPTY = (1 << 64) - 1


class MgpuPlan(C.Structure):
    """== hd_mgpu_plan (include/hd_b200.h)"""
    _fields_ = [(n, C.c_uint64) for n in ("seg_begin", "seg_end", "first_unit", "halo_units", "shard_units",
                                          "readable_units", "first_gap_word", "n_gap_words", "out_capacity",
                                          "workspace_bytes")]


def _lib():
    from . import lib
    return lib()


def plan_shard(scheme, n_units, rank, world, min_code_length=1, halo_segments=0):
    """hd_mgpu_plan_shard: what rank `rank` of `world` uploads and decodes."""
    from . import _check
    p = MgpuPlan()
    _check(_lib().hd_mgpu_plan_shard(scheme, n_units, rank, world, min_code_length, halo_segments, C.byref(p)),
           "hd_mgpu_plan_shard")
    return p


def plan_shards(n_units, world, halo_segments=HALO_SEGMENTS):
    """[(seg_begin, seg_end, halo_segments)] per rank of a self-sync decode (hd_mgpu_plan_shard)."""
    res = []
    for r in range(world):
        p = plan_shard(SCHEME_SELFSYNC, n_units, r, world, 1, halo_segments)
        res.append((int(p.seg_begin), int(p.seg_end), int(p.halo_units) // SEGMENT_UNITS))
    return res


def _infos_array(infos):
    rows = [[int(v) & EMPTY for v in row] for row in np.asarray(infos, dtype=object).reshape(-1, 3)]
    return np.ascontiguousarray(np.array(rows, dtype=np.uint64))


def place_outputs(counts, n_symbols):
    """hd_mgpu_place for every rank: [(offset, valid_count)], clamped to the stream's symbol count."""
    world = len(counts)
    infos = np.zeros((world, 3), dtype=np.uint64)
    infos[:, 0] = np.array([int(c) for c in counts], dtype=np.uint64)
    res = []
    off, cnt = C.c_uint64(0), C.c_uint64(0)
    for r in range(world):
        _lib().hd_mgpu_place(infos.ctypes.data, world, int(n_symbols), r, C.byref(off), C.byref(cnt))
        res.append((int(off.value), int(cnt.value)))
    return res


def seams_ok(infos):
    """Ranks whose assumed entry offset was wrong given their predecessor's exit (repeated hd_mgpu_first_bad_seam)."""
    a = _infos_array(infos)
    world = len(a)
    bad = []
    while True:
        g = int(_lib().hd_mgpu_first_bad_seam(a.ctypes.data, world))
        if g >= world:
            return bad
        bad.append(g)
        # pretend the seam is repaired (entry := what the predecessor left) and look for the next one
        prev = [k for k in range(g) if int(a[k, 1]) != EMPTY]
        a[g, 1] = a[prev[-1], 2]


def resolve_seams(rank, world, infos, gather, redecode, n_symbols):
    """The seam protocol of hd_mgpu_finish with a pluggable decode (host-logic tests on CPU / gloo): infos = gathered
    (count, entry, exit) rows; gather() re-gathers after a repair; redecode(first_bit) decodes THIS rank's shard again
    from an exact entry offset.  A wrong seam can only be repaired once its predecessor is right, so at most world-1
    rounds (sequential chain: the worst case of a code that never self-synchronises)."""
    L = _lib()
    for _ in range(world + 1):
        a = _infos_array(infos)
        g = int(L.hd_mgpu_first_bad_seam(a.ctypes.data, world))
        if g >= world:
            return place_outputs([int(row[0]) for row in a], n_symbols)
        if rank == g:
            prev = [k for k in range(g) if int(a[k, 1]) != EMPTY]
            redecode(int(a[prev[-1], 2]) & 15)
        infos = gather()
    raise RuntimeError("shard seams did not converge")


class _ShardedDecoder:
    """One rank of a sharded decode: device buffers + an hd_mgpu context.  Subclasses upload their inputs."""
    scheme = SCHEME_SELFSYNC

    def _setup(self, n_units, table, n_symbols, rank, world, dist, use_mailbox, halo_segments=0, allgather=None):
        import torch
        from . import lib, _check, _stream_ptr
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self._lib, self._check, self._stream_ptr = lib(), _check, _stream_ptr
        self.table = table
        self.n_symbols = int(n_symbols)
        self.plan = plan_shard(self.scheme, n_units, rank, world, max(1, table.min_length), halo_segments)
        p = self.plan
        self.n_units_local = int(p.shard_units)
        self.out_cap = max(64, int(p.out_capacity))
        self.out = torch.empty(self.out_cap, dtype=torch.uint8, device="cuda")
        self.ws = torch.empty(int(p.workspace_bytes) + 128, dtype=torch.uint8, device="cuda")
        self.ws_ptr = self.ws.data_ptr() + (-self.ws.data_ptr()) % 64
        self.count_local = 0
        self.offset = 0
        self.ctx = C.c_void_p(None)
        self._check(self._lib.hd_mgpu_create(self.scheme, rank, world, C.byref(p), C.byref(self.ctx)), "hd_mgpu_create")
        self.mailbox = None
        self._cb = None
        if world > 1:
            if allgather is not None:      # caller-supplied exchange: allgather(d_send, d_recv, nbytes, stream, user) -> 0
                self._cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)(allgather)
                self._check(self._lib.hd_mgpu_exchange_callback(self.ctx, self._cb, None), "hd_mgpu_exchange_callback")
                return
            if dist is not None and use_mailbox:
                self._setup_mailbox()
            if self.mailbox is not None:
                self._check(self._lib.hd_mgpu_exchange_mailbox(self.ctx, self.mailbox), "hd_mgpu_exchange_mailbox")
            else:
                self._setup_callback()

    def _setup_mailbox(self):
        """Peer-memory mailbox over NVLink (hd_mailbox_*): handles are exchanged once with torch.distributed."""
        torch, L = self.torch, self._lib
        h = C.c_void_p(None)
        ok = L.hd_mailbox_create(self.world, self.rank, C.byref(h)) == 0
        handle = np.zeros(64, dtype=np.uint8)
        if ok:
            ok = L.hd_mailbox_export(h, handle.ctypes.data) == 0
        mine = torch.from_numpy(handle).cuda()
        every = torch.zeros(64 * self.world, dtype=torch.uint8, device="cuda")
        self.dist.all_gather_into_tensor(every, mine)
        if ok:
            handles = np.ascontiguousarray(every.cpu().numpy())
            ok = L.hd_mailbox_connect(h, handles.ctypes.data) == 0
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)      # all ranks or none
        if int(t.item()):
            self.mailbox = h
        elif h:
            L.hd_mailbox_destroy(h)

    def _setup_callback(self):
        """Fallback exchange when CUDA IPC is not available: torch.distributed's (NCCL) all-gather of the 24 bytes,
        handed to the library as the hd_allgather_fn callback; it works in place on the library's device buffers
        (zero-copy tensor views through __cuda_array_interface__)."""
        torch, dist, world = self.torch, self.dist, self.world
        views = {}

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}

        def allgather(d_send, d_recv, nbytes, stream, user):
            try:
                key = (d_send, d_recv)
                if key not in views:
                    views[key] = (torch.as_tensor(_Dev(d_send, nbytes // 8), device="cuda"),
                                  torch.as_tensor(_Dev(d_recv, nbytes // 8 * world), device="cuda"))
                send, recv = views[key]
                dist.all_gather_into_tensor(recv, send)
                return 0
            except Exception:
                return 1
        self._cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)(allgather)
        self._check(self._lib.hd_mgpu_exchange_callback(self.ctx, self._cb, None), "hd_mgpu_exchange_callback")

    # ---- the step
    def step(self):
        """The timed step, entirely on the stream (hd_mgpu_enqueue): local shard decode + the exchange of {count,
        entry, exit}.  Several steps may be enqueued back to back; call wait() / finish() before using the results."""
        self._check(self._lib.hd_mgpu_enqueue(self.ctx, self._stream_ptr(self.torch)), "hd_mgpu_enqueue")

    def align(self):
        """Stream-side rendezvous of the ranks (hd_mgpu_align: one exchange round whose result nobody reads): used by
        bench.py after the untimed L2 flush, so that a rank whose flush ran longer is not waited for inside the next
        timed step."""
        if self.world > 1:
            self._check(self._lib.hd_mgpu_align(self.ctx, self._stream_ptr(self.torch)), "hd_mgpu_align")

    def finish(self):
        """Collective host half (hd_mgpu_finish): convergence, seams, repair, output placement of this rank."""
        off, cnt = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.hd_mgpu_finish(self.ctx, self.n_symbols, C.byref(off), C.byref(cnt),
                                             self._stream_ptr(self.torch)), "hd_mgpu_finish")
        self.offset, self.count_local = int(off.value), int(cnt.value)
        return self.placed()

    wait = finish

    def infos(self):
        a = np.zeros((self.world, 3), dtype=np.uint64)
        self._check(self._lib.hd_mgpu_infos(self.ctx, a.ctypes.data), "hd_mgpu_infos")
        return a

    def repairs(self):
        """decodes this rank repeated in the last finish() (hd_mgpu_repair_count)"""
        return int(self._lib.hd_mgpu_repair_count(self.ctx))

    def placed(self):
        return place_outputs([int(c) for c in self.infos()[:, 0]], self.n_symbols)

    def verify(self, data):
        self.step()
        placed = self.finish()
        want = self.torch.from_numpy(np.ascontiguousarray(data[self.offset:self.offset + self.count_local])).cuda()
        total = sum(c for _, c in placed)
        return total == self.n_symbols and bool(self.torch.equal(self.out[:self.count_local], want))

    def gather_output(self, dst=0):
        """Optional last step: collect the sharded output on rank `dst` over NVLink.  finish() must have
        placed the outputs.  Point-to-point NCCL sends of each rank's bytes straight into their place in the
        destination buffer; returns the whole output on `dst`, None elsewhere."""
        torch, dist = self.torch, self.dist
        placed = self.placed()
        if self.world == 1 or dist is None:
            return self.out[:self.count_local]
        if self.rank == dst:
            whole = torch.empty(self.n_symbols, dtype=torch.uint8, device="cuda")
            off, cnt = placed[dst]
            whole[off:off + cnt].copy_(self.out[:cnt])
            reqs = [dist.irecv(whole[o:o + c], src=r) for r, (o, c) in enumerate(placed) if r != dst and c]
            for q in reqs:
                q.wait()
            return whole
        if self.count_local:
            dist.send(self.out[:self.count_local], dst=dst)
        return None

    def close(self):
        if self.ctx:
            self.torch.cuda.synchronize()
            if self.dist is not None and self.world > 1:
                self.dist.barrier()          # nobody may still be writing into a mailbox that is about to go
            self._lib.hd_mgpu_destroy(self.ctx)
            self.ctx = C.c_void_p(None)
            if self.mailbox is not None:
                self._lib.hd_mailbox_destroy(self.mailbox)
                self.mailbox = None

    # ---- end to end from / to pinned host memory
    def _h_gaps_ptr(self):
        return None

    def e2e(self, data, steps=5):
        """Host-buffer path (hd_mgpu_host_step): pinned shard -> H2D -> decode -> D2H as a three-stage pipeline over
        pieces on this rank's GPU (both PCIe directions busy at once), then the exchange of the counts and the
        collective finish.  Wall clock per step, max over ranks."""
        import time
        from . import HostDecoder
        torch, p = self.torch, self.plan
        self.finish()
        cap = max(64, int(p.out_capacity))
        h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
        hdec = HostDecoder(max(4, int(p.readable_units)), cap, device=torch.cuda.current_device())
        sp = self._stream_ptr(torch)
        tot = 0.0
        for i in range(steps + 1):
            if self.dist is not None and self.world > 1:
                self.dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            self._check(self._lib.hd_mgpu_host_step(self.ctx, hdec.handle, C.c_void_p(self.h_in.data_ptr()), self._h_gaps_ptr(),
                                                    C.c_void_p(h_out.data_ptr()), cap, sp), "hd_mgpu_host_step")
            self.finish()
            if self.repairs():
                self._check(self._lib.hd_mgpu_host_refresh(self.ctx, C.c_void_p(h_out.data_ptr()), self.count_local, sp),
                            "hd_mgpu_host_refresh")
            if i:
                tot += time.perf_counter() - t0
        t = torch.tensor([tot / steps], dtype=torch.float64, device="cuda")
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        got, want = h_out.numpy()[:self.count_local], data[self.offset:self.offset + self.count_local]
        last = self.offset + self.count_local == self.n_symbols     # (reference encoder defect D1: last symbol)
        ok = bool(np.array_equal(got, want)) or (last and self.count_local > 1 and bool(np.array_equal(got[:-1], want[:-1])))
        # back to the decoder's own device buffers for further device-side steps
        self._bind()
        hdec.close()
        if not ok:
            raise RuntimeError("host-buffer shard decode mismatch")
        h2d = torch.tensor([self._h2d_bytes()], dtype=torch.int64, device="cuda")
        d2h = torch.tensor([self.count_local], dtype=torch.int64, device="cuda")
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(h2d)
            self.dist.all_reduce(d2h)
        return {"value": round(self.n_symbols / float(t.item()) / 1e9, 3), "unit": "GB/s",
                "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item()), "steps": steps,
                "api": "hd_mgpu_host_step + hd_mgpu_finish per rank (pinned host shard in/out, pipelined H2D / decode / "
                       "D2H, wall clock, max over ranks)"}


class ShardedSelfSyncDecoder(_ShardedDecoder):
    """One rank's shard of a self-synchronising decode (CUDA, hd_mgpu_* with HD_SCHEME_SELFSYNC)."""
    scheme = SCHEME_SELFSYNC

    def __init__(self, units, table, n_symbols, rank, world, dist=None, halo_segments=HALO_SEGMENTS, use_mailbox=True,
                 allgather=None):
        self._setup(len(units), table, n_symbols, rank, world, dist, use_mailbox, halo_segments, allgather)
        torch, p = self.torch, self.plan
        u0, n = int(p.first_unit), max(4, int(p.readable_units))
        self.h_units = np.zeros(n, dtype=np.uint32)
        self.h_units[:int(p.readable_units)] = units[u0:u0 + int(p.readable_units)]   # private copy of the shard
        self.h_in = torch.from_numpy(self.h_units.view(np.int32)).pin_memory()
        self.d_units = self.h_in.cuda()
        self._bind()

    def _bind(self):
        self._check(self._lib.hd_mgpu_bind(self.ctx, C.c_void_p(self.d_units.data_ptr()), None, self.table.handle,
                                           C.c_void_p(self.out.data_ptr()), C.c_void_p(self.ws_ptr)), "hd_mgpu_bind")

    def _h2d_bytes(self):
        return self.h_in.numel() * 4


class ShardedGapDecoder(_ShardedDecoder):
    """One rank's shard of a gap-array decode (hd_mgpu_* with HD_SCHEME_GAP): the segment entry rule of
    decoder.cu:505-507 makes every 128-bit segment boundary a valid entry point, so shards need no halo and no repair;
    only the symbol counts are exchanged to place the outputs."""
    scheme = SCHEME_GAP

    def __init__(self, units, gap_words, table, n_symbols, rank, world, dist=None, use_mailbox=True, allgather=None):
        self._setup(len(units), table, n_symbols, rank, world, dist, use_mailbox, 0, allgather)
        torch, p = self.torch, self.plan
        u0, n = int(p.first_unit), max(4, int(p.readable_units))
        self.h_units = np.zeros(n, dtype=np.uint32)
        self.h_units[:int(p.readable_units)] = units[u0:u0 + int(p.readable_units)]
        g0, ng = int(p.first_gap_word), max(1, int(p.n_gap_words))
        self.h_gaps = np.zeros(ng, dtype=np.uint32)
        self.h_gaps[:int(p.n_gap_words)] = gap_words[g0:g0 + int(p.n_gap_words)]
        self.h_in = torch.from_numpy(self.h_units.view(np.int32)).pin_memory()
        self.h_gin = torch.from_numpy(self.h_gaps.view(np.int32)).pin_memory()
        self.d_units = self.h_in.cuda()
        self.d_gaps = self.h_gin.cuda()
        self._bind()

    def _bind(self):
        self._check(self._lib.hd_mgpu_bind(self.ctx, C.c_void_p(self.d_units.data_ptr()), C.c_void_p(self.d_gaps.data_ptr()),
                                           self.table.handle, C.c_void_p(self.out.data_ptr()), C.c_void_p(self.ws_ptr)),
                    "hd_mgpu_bind")

    def _h_gaps_ptr(self):
        return C.c_void_p(self.h_gin.data_ptr())

    def _h2d_bytes(self):
        return (self.h_in.numel() + self.h_gin.numel()) * 4
