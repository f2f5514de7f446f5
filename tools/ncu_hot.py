"""Per-SASS-instruction samples / executed counts from an .ncu-rep source page (SASS view), top-N and phase totals."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
c = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append(dict(addr=r[c["Address"]], sass=r[c["Source"]].strip(), samples=int(r[c["# Samples"]] or 0),
                    execd=int(r[c["Instructions Executed"]] or 0), thr=float(r[c["Avg. Threads Executed"]] or 0),
                    wf=int(r[c["L1 Wavefronts Shared"]] or 0), bar=int(r[c["stall_barrier"]] or 0),
                    lsb=int(r[c["stall_long_sb"]] or 0), ssb=int(r[c["stall_short_sb"]] or 0),
                    wait=int(r[c["stall_wait"]] or 0), math=int(r[c["stall_math"]] or 0)))
tot_s = sum(i["samples"] for i in ins)
tot_e = sum(i["execd"] for i in ins)
print("instructions %d  total samples %d  total executed %d" % (len(ins), tot_s, tot_e))
print("idx  samples%  exec%  thr  wf      bar  lsb  ssb wait math  sass")
for k, i in enumerate(ins):
    i["idx"] = k
for i in sorted(ins, key=lambda x: -x["samples"])[:top]:
    print("%4d %6.2f %6.2f %5.1f %8d %5d %5d %5d %5d %5d  %s" % (i["idx"], 100.0 * i["samples"] / tot_s, 100.0 * i["execd"] / tot_e,
          i["thr"], i["wf"], i["bar"], i["lsb"], i["ssb"], i["wait"], i["math"], i["sass"][:70]))
if len(sys.argv) > 3:   # dump a range idx0:idx1
    a, b = map(int, sys.argv[3].split(":"))
    for i in ins[a:b]:
        print("%4d %6.2f %6.2f %5.1f %8d  %s" % (i["idx"], 100.0 * i["samples"] / tot_s, 100.0 * i["execd"] / tot_e, i["thr"], i["wf"], i["sass"][:90]))
