"""Per-source-line totals sorted by stall samples, plus sums over line ranges: ncu_lines2.py rep kernel [file:lo-hi ...]"""
import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No":
        hdr = r; c = {}
        for i, h in enumerate(r): c.setdefault(h, i)
        continue
    if hdr and len(r) >= len(hdr):
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[c["# Samples"]] or 0), int(r[c["Instructions Executed"]] or 0)))
        except ValueError:
            pass
ts = sum(l[3] for l in lines); te = sum(l[4] for l in lines)
print("samples %d executed %d" % (ts, te))
for spec in sys.argv[3:]:
    f, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
    s = sum(l[3] for l in lines if l[0] == f and lo <= l[1] <= hi); e = sum(l[4] for l in lines if l[0] == f and lo <= l[1] <= hi)
    print("%s  samp%% %5.2f  exec%% %5.2f" % (spec, 100.0 * s / ts, 100.0 * e / te))
if len(sys.argv) <= 3:
    for l in sorted(lines, key=lambda x: -x[3])[:50]:
        print("%-16s %4d  samp%% %5.2f exec%% %5.2f  %s" % (l[0], l[1], 100.0 * l[3] / ts, 100.0 * l[4] / te, l[2][:100]))
