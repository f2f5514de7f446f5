"""Hottest SASS instructions per kernel of an .ncu-rep (source page) -> markdown (profiles/r1_ncu_hot_instructions.md)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print("# Hottest SASS instructions per kernel (ncu --set full --import-source on, self-sync 100 MiB r=0.5, B200)\n")
print("samples% = share of warp-state samples at this instruction; exec% = share of executed warp instructions; thr = avg active threads\n")
name = None; hdr = None; ins = []
def flush():
    if not ins: return
    ts = sum(i["s"] for i in ins) or 1; te = sum(i["e"] for i in ins) or 1
    print("## %s\n\n%d SASS instructions, %d samples, %d executed warp instructions\n" % (name, len(ins), ts, te))
    print("| samples% | exec% | thr | smem wavefronts | barrier | long_sb | short_sb | math | wait | SASS |\n|---|---|---|---|---|---|---|---|---|---|")
    for i in sorted(ins, key=lambda x: -x["s"])[:top]:
        print("| %.2f | %.2f | %.1f | %d | %d | %d | %d | %d | %d | `%s` |" % (100.0 * i["s"] / ts, 100.0 * i["e"] / te, i["t"], i["wf"], i["bar"], i["lsb"], i["ssb"], i["math"], i["wait"], i["sass"][:60]))
    print()
for r in rows:
    if r and r[0] in ("Kernel Name", "Function Name"):
        flush(); ins = []; name = r[1]; continue
    if "Address" in r and "Source" in r:
        hdr = r; c = {h: i for i, h in enumerate(hdr)}; continue
    if hdr and len(r) >= len(hdr):
        try:
            ins.append(dict(sass=r[c["Source"]].strip(), s=int(r[c["# Samples"]] or 0), e=int(r[c["Instructions Executed"]] or 0),
                            t=float(r[c["Avg. Threads Executed"]] or 0), wf=int(r[c["L1 Wavefronts Shared"]] or 0),
                            bar=int(r[c["stall_barrier"]] or 0), lsb=int(r[c["stall_long_sb"]] or 0), ssb=int(r[c["stall_short_sb"]] or 0),
                            wait=int(r[c["stall_wait"]] or 0), math=int(r[c["stall_math"]] or 0)))
        except ValueError:
            pass
flush()
