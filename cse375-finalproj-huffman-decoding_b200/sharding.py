"""Multi-GPU sharding of ONE Huffman stream (one process per GPU, torch.distributed for the plumbing).

Replaces gpuhd/multigpu_demo.cc:178-227,255-313, which slices the compressed units AND the symbols into equal 1/n
parts (wrong by construction: the reference README says "MULTI-GPU BUT TESTS FAIL", README.md:37).  Here the stream
is cut on 128-bit subsequence boundaries (whole 256-segment tiles); every rank decodes its shard independently from
local HBM and the only exchange is an all-gather of three integers per rank:

    {symbols decoded, entry bit offset assumed, bit offset at which the parse leaves the shard}

entry[g] must equal exit[g-1] (it does once the left halo is longer than the self-synchronisation distance; a
mismatching rank re-decodes from the exact offset), and the exclusive scan of the counts places the outputs.
The decode itself is pluggable so that the host logic is testable on CPU with the gloo backend.
"""
import ctypes as C

import numpy as np

HALO_SEGMENTS = 4096          # left halo: 4096 x 128 bit = 64 KiB of stream (>> observed max sync distance ~5 Kbit)
SEGMENT_UNITS = 4
TILE_SEGMENTS = 256


def plan_shards(n_units, world, halo_segments=HALO_SEGMENTS):
    """[(seg_begin, seg_end, halo_segments)] per rank; boundaries are whole tiles (hd_shard_segments)."""
    n_seg = (n_units + SEGMENT_UNITS - 1) // SEGMENT_UNITS
    n_tiles = (n_seg + TILE_SEGMENTS - 1) // TILE_SEGMENTS
    plan = []
    for r in range(world):
        b = (n_tiles * r // world) * TILE_SEGMENTS
        e = min((n_tiles * (r + 1) // world) * TILE_SEGMENTS, n_seg)
        h = min(halo_segments, b) if r > 0 else 0
        h -= h % 8
        plan.append((b, e, h))
    return plan


def place_outputs(counts, n_symbols):
    """Exclusive scan of per-rank counts -> (offset, valid_count) per rank, clamped to the stream's symbol count
    (the last shard also decodes the pad bits)."""
    res, off = [], 0
    for c in counts:
        c = int(c)
        valid = max(0, min(c, n_symbols - off))
        res.append((off, valid))
        off += c
    return res


def seams_ok(infos):
    """infos[g] = (count, entry, exit).  Returns the list of ranks whose assumed entry offset was wrong."""
    bad = []
    for g in range(1, len(infos)):
        if int(infos[g][1]) != int(infos[g - 1][2]):
            bad.append(g)
    return bad


def resolve_seams(rank, world, infos, gather, redecode, n_symbols):
    """The seam protocol, independent of where the decode runs.  infos = gathered (count, entry, exit) rows;
    gather() re-gathers after a repair; redecode(first_bit) re-decodes THIS rank's shard from an exact entry offset.
    A wrong seam can only be repaired once its predecessor is right, so at most world-1 rounds (sequential chain,
    the worst case of a code that never self-synchronises)."""
    for _ in range(world):
        bad = seams_ok(infos)
        if not bad:
            return place_outputs([row[0] for row in infos], n_symbols)
        g = bad[0]
        if rank == g:
            redecode(int(infos[g - 1][2]))
        infos = gather()
    raise RuntimeError("shard seams did not converge")


class ShardedSelfSyncDecoder:
    """One rank's shard of a self-synchronising decode (CUDA)."""


    def __init__(self, units, table, n_symbols, rank, world, dist=None, halo_segments=HALO_SEGMENTS,
                 use_mailbox=True):
        import torch
        from . import lib, _check, _stream_ptr
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self._lib, self._check, self._stream_ptr = lib(), _check, _stream_ptr
        self.table = table
        self.n_symbols = int(n_symbols)
        n_units = len(units)
        self.plan = plan_shards(n_units, world, halo_segments)
        b, e, h = self.plan[rank]
        self.seg_begin, self.seg_end, self.halo = b, e, h
        u0 = (b - h) * SEGMENT_UNITS
        u1 = min(e * SEGMENT_UNITS, n_units)
        u_read = min(u1 + SEGMENT_UNITS, n_units)            # right halo: the last codeword may end there
        self.n_units_halo = h * SEGMENT_UNITS
        self.n_units_local = u1 - b * SEGMENT_UNITS
        self.n_units_readable = u_read - u0
        self.h_units = np.array(units[u0:u_read], dtype=np.uint32)   # private, writable copy of the shard
        self.d_units = torch.from_numpy(self.h_units.view(np.int32)).cuda()
        min_len = max(1, table.min_length)
        self.out_cap = (self.n_units_local * 32 + min_len - 1) // min_len + 64
        self.out = torch.empty(self.out_cap, dtype=torch.uint8, device="cuda")
        nbytes = self._lib.hd_workspace_bytes_selfsync(self.n_units_halo + self.n_units_local)
        self.ws = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
        self.ws_ptr = self.ws.data_ptr() + (-self.ws.data_ptr()) % 64
        self.ws_bytes = nbytes
        self.info = torch.zeros(3, dtype=torch.int64, device="cuda")
        self.all_info = torch.zeros(world * 3, dtype=torch.int64, device="cuda")
        self.count_local = 0
        self.offset = 0
        self._flag = None
        self._outstanding = []
        self._next_slot = 0
        # the per-step exchange: peer-memory mailbox over NVLink (hd_mailbox_*) if CUDA IPC works between all ranks,
        # else an NCCL all-gather of the same 24 bytes
        self.mailbox = None
        self.seq = 0
        if dist is not None and world > 1 and use_mailbox:
            self._setup_mailbox()

    def _setup_mailbox(self):
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

    def _exchange(self):
        """All-gather of self.info into self.all_info, on the stream (collective: every rank calls it)."""
        if self.mailbox is not None:
            self.seq += 1
            sp = self._stream_ptr(self.torch)
            self._check(self._lib.hd_mailbox_publish(self.mailbox, C.c_void_p(self.info.data_ptr()), self.seq, sp),
                        "hd_mailbox_publish")
            self._check(self._lib.hd_mailbox_wait(self.mailbox, self.seq, C.c_void_p(self.all_info.data_ptr()), sp),
                        "hd_mailbox_wait")
        elif self.dist is not None and self.world > 1:
            self.dist.all_gather_into_tensor(self.all_info, self.info)
        else:
            self.all_info.copy_(self.info)

    def align(self):
        """Stream-side rendezvous of the ranks (one exchange round whose result nobody reads): used by bench.py after
        the untimed L2 flush, so that a rank whose flush ran longer is not waited for inside the next timed step."""
        if self.world > 1:
            self._exchange()

    def close(self):
        if self.mailbox is not None:
            self.torch.cuda.synchronize()
            if self.dist is not None:
                self.dist.barrier()          # nobody may still be writing into a mailbox that is about to go
            self._lib.hd_mailbox_destroy(self.mailbox)
            self.mailbox = None

    def _args(self, first_bit, halo):
        hu = self.n_units_halo if halo else 0
        ptr = self.d_units.data_ptr() + 4 * (self.n_units_halo - hu)
        return (C.c_void_p(ptr), hu, self.n_units_local, self.n_units_readable - (self.n_units_halo - hu), first_bit,
                self.table.handle, C.c_void_p(self.out.data_ptr()), self.out_cap, C.c_void_p(self.ws_ptr), self.ws_bytes,
                C.c_void_p(self.info.data_ptr()))

    def _decode(self, first_bit=0, halo=True):
        """Synchronous shard decode (hd_decode_selfsync_shard)."""
        self._check(self._lib.hd_decode_selfsync_shard(*self._args(first_bit, halo), self._stream_ptr(self.torch)),
                    "hd_decode_selfsync_shard")

    def step(self):
        """The timed step, entirely on the stream: local shard decode (hd_decode_selfsync_shard_enqueue) + the
        exchange of {count, entry, exit}.  Several steps may be enqueued back to back (each has its own convergence
        flag, up to 256 outstanding); call wait() before using the results."""
        if self._flag is None:
            self._flag = self.torch.zeros(256, dtype=self.torch.int32).pin_memory()
        if len(self._outstanding) >= 256:
            self.wait()
        slot = self._next_slot
        self._next_slot = (slot + 1) % 256
        self._outstanding.append(slot)
        args = self._args(0, True) + (C.c_void_p(self._flag.data_ptr() + 4 * slot), self._stream_ptr(self.torch))
        self._check(self._lib.hd_decode_selfsync_shard_enqueue(*args), "hd_decode_selfsync_shard_enqueue")
        self._exchange()

    def wait(self):
        """Host half of the enqueued steps: synchronise the stream and look at their convergence flags.  In the rare
        unconverged case (a parse that did not synchronise within 64 Kbit) the shard is decoded again synchronously
        (iterating the synchronisation to its fixed point) and the counts are exchanged again."""
        if not self._outstanding:
            return
        self.torch.cuda.current_stream().synchronize()
        redone = 0
        for slot in self._outstanding:
            if int(self._flag[slot].item()):
                redone = 1
        self._outstanding = []
        if redone:
            self._decode()
        # every rank must take part in a re-exchange if ANY rank had to decode again
        if self.dist is not None and self.world > 1:
            t = self.torch.tensor([redone], dtype=self.torch.int64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            if int(t.item()):
                self._exchange()
                self.torch.cuda.current_stream().synchronize()
        elif redone:
            self.all_info.copy_(self.info)
        if self.mailbox is not None and self._lib.hd_mailbox_timed_out(self.mailbox):
            raise RuntimeError("shard exchange timed out: a peer rank never published its counts")

    def _gather(self):
        self._exchange()
        return self.all_info.cpu().numpy().reshape(self.world, 3)

    def finish(self):
        """Host side of the exchange: check the seams, repair if a halo was too short, place the outputs."""
        self.wait()
        infos = self.all_info.cpu().numpy().reshape(self.world, 3)
        placed = resolve_seams(self.rank, self.world, infos, self._gather,
                               lambda first_bit: self._decode(first_bit=first_bit, halo=False), self.n_symbols)
        self.offset, self.count_local = placed[self.rank]
        return placed

    def verify(self, data):
        self.step()
        self.finish()
        want = self.torch.from_numpy(np.ascontiguousarray(data[self.offset:self.offset + self.count_local])).cuda()
        total = sum(c for _, c in place_outputs(self.all_info.cpu().numpy().reshape(self.world, 3)[:, 0],
                                                self.n_symbols))
        return total == self.n_symbols and bool(self.torch.equal(self.out[:self.count_local], want))

    def gather_output(self, dst=0):
        """Optional last step: collect the sharded output on rank `dst` over NVLink.  finish() must have
        placed the outputs (every rank knows all counts and offsets).  Point-to-point NCCL sends of each rank's bytes
        straight into their place in the destination buffer; returns the whole output on `dst`, None elsewhere."""
        torch, dist = self.torch, self.dist
        infos = self.all_info.cpu().numpy().reshape(self.world, 3)
        placed = place_outputs(infos[:, 0], self.n_symbols)
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

    def e2e(self, data, steps=5):
        """Host-buffer path: pinned shard -> H2D -> decode + all-gather -> D2H, wall clock, max over ranks."""
        import time
        torch = self.torch
        h_in = torch.from_numpy(self.h_units.view(np.int32)).pin_memory()
        h_out = torch.empty(self.count_local, dtype=torch.uint8).pin_memory()
        tot = 0.0
        for i in range(steps + 1):
            if self.dist is not None:
                self.dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            self.d_units.copy_(h_in, non_blocking=True)
            self.step()
            self.wait()
            h_out.copy_(self.out[:self.count_local], non_blocking=True)
            torch.cuda.synchronize()
            if i:
                tot += time.perf_counter() - t0
        t = torch.tensor([tot / steps], dtype=torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        ok = bool(np.array_equal(h_out.numpy(), data[self.offset:self.offset + self.count_local]))
        if not ok:
            raise RuntimeError("host-buffer shard decode mismatch")
        h2d = torch.tensor([h_in.numel() * 4], dtype=torch.int64, device="cuda")
        d2h = torch.tensor([h_out.numel()], dtype=torch.int64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(h2d)
            self.dist.all_reduce(d2h)
        return {"value": round(self.n_symbols / float(t.item()) / 1e9, 3), "unit": "GB/s",
                "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item()), "steps": steps,
                "api": "ShardedSelfSyncDecoder (pinned host shard in/out, wall clock, max over ranks)"}
