"""In-stream cost of one hd_mailbox_publish + hd_mailbox_wait pair on one GPU (world 1: no peer skew)."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_b200 as hd

torch.cuda.set_device(0)
L = hd.lib()
h = C.c_void_p(None)
assert L.hd_mailbox_create(1, 0, C.byref(h)) == 0
handle = np.zeros(64, dtype=np.uint8)
assert L.hd_mailbox_export(h, handle.ctypes.data) == 0
assert L.hd_mailbox_connect(h, handle.ctypes.data) == 0
info = torch.tensor([1, 2, 3], dtype=torch.int64, device="cuda")
allinfo = torch.zeros(3, dtype=torch.int64, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.zeros(64 << 20, device="cuda")   # one add_ = 0.1 ms: the host runs ahead of the GPU, as in the bench
seq = 1
for reps in (1, 1, 1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.add_(1); x.add_(1)
    a.record()
    for _ in range(50):
        x.add_(1)          # stands for the decode kernels (keeps the stream busy so launch latency is hidden as in the bench)
        L.hd_mailbox_publish(h, C.c_void_p(info.data_ptr()), seq, sp)
        L.hd_mailbox_wait(h, seq, C.c_void_p(allinfo.data_ptr()), sp)
        seq += 1
    b.record(); torch.cuda.synchronize()
    t1 = a.elapsed_time(b) / 50
    a.record()
    for _ in range(50):
        x.add_(1)
    b.record(); torch.cuda.synchronize()
    t0 = a.elapsed_time(b) / 50
    print("per step: with publish+wait %.2f us, without %.2f us -> pair costs %.2f us" % (t1 * 1e3, t0 * 1e3, (t1 - t0) * 1e3))
L.hd_mailbox_destroy(h)
