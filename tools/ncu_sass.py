"""Per-SASS-instruction view of ONE kernel of an .ncu-rep: top-N by samples, or a dump of an index range.
usage: ncu_sass.py <rep> <kernel regex> [top N | a:b]"""
import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
arg = sys.argv[3] if len(sys.argv) > 3 else "40"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
c = {}
for i, h in enumerate(hdr):
    c.setdefault(h, i)
ins = []
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    g = lambda k: int(r[c[k]] or 0)
    ins.append(dict(sass=r[c["Source"]].strip(), samples=g("# Samples"), execd=g("Instructions Executed"),
                    thr=float(r[c["Avg. Threads Executed"]] or 0), wf=g("L1 Wavefronts Shared"), lsb=g("stall_long_sb"),
                    ssb=g("stall_short_sb"), wait=g("stall_wait"), math=g("stall_math"), bar=g("stall_barrier"),
                    nsel=g("stall_not_selected"), mio=g("stall_mio"), br=g("stall_branch_resolving")))
ts = sum(i["samples"] for i in ins) or 1
te = sum(i["execd"] for i in ins) or 1
print("instructions %d  samples %d  executed %d" % (len(ins), ts, te))
fmt = "%4d samp%%%6.2f exec%%%6.2f thr%5.1f wf%8d | lsb%4d ssb%4d wait%4d math%4d bar%4d nsel%4d mio%4d br%4d | %s"
def show(k, i):
    print(fmt % (k, 100.0 * i["samples"] / ts, 100.0 * i["execd"] / te, i["thr"], i["wf"], i["lsb"], i["ssb"], i["wait"], i["math"],
                 i["bar"], i["nsel"], i["mio"], i["br"], i["sass"][:64]))
if ":" in arg:
    a, b = map(int, arg.split(":"))
    for k in range(a, min(b, len(ins))):
        show(k, ins[k])
else:
    for k, i in sorted(enumerate(ins), key=lambda x: -x[1]["samples"])[:int(arg)]:
        show(k, i)
