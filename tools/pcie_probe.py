"""PCIe floor for the end-to-end path: pinned H2D of the compressed stream and D2H of the decoded bytes, alone and together."""
import time
import torch

torch.cuda.set_device(0)
n_in, n_out = 78 << 20, 100 << 20
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2]


for _ in range(2):
    run(True, True)
t = run(True, False); print("H2D alone  %.3f ms  %.1f GB/s" % (t * 1e3, n_in / t / 1e9))
t = run(False, True); print("D2H alone  %.3f ms  %.1f GB/s" % (t * 1e3, n_out / t / 1e9))
t = run(True, True); print("both       %.3f ms  H2D %.1f GB/s + D2H %.1f GB/s  (decoded bytes / time = %.1f GB/s)" % (t * 1e3, n_in / t / 1e9, n_out / t / 1e9, n_out / t / 1e9))


def run_pieces(k, reps=10, lag=0):
    """both directions in k pieces each; D2H piece i may only start after H2D piece i+lag-1 (event), like the pipeline"""
    ts = []
    ci, co = n_in // k, n_out // k
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        evs = []
        with torch.cuda.stream(s1):
            for i in range(k):
                d_in[i * ci:(i + 1) * ci].copy_(h_in[i * ci:(i + 1) * ci], non_blocking=True)
                e = torch.cuda.Event(); e.record(s1); evs.append(e)
        with torch.cuda.stream(s2):
            for i in range(k):
                if lag:
                    s2.wait_event(evs[i])
                h_out[i * co:(i + 1) * co].copy_(d_out[i * co:(i + 1) * co], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2]


for k in (1, 5, 10, 20):
    t = run_pieces(k); t2 = run_pieces(k, lag=1)
    print("%2d pieces per direction: independent %.3f ms (%.1f GB/s decoded), D2H(i) after H2D(i) %.3f ms (%.1f GB/s)" % (
        k, t * 1e3, n_out / t / 1e9, t2 * 1e3, n_out / t2 / 1e9))
