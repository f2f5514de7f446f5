"""Per-source-line instruction and sample totals of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; c = {}
    if r and r[0] == "Line No":
        for i, h in enumerate(r): c.setdefault(h, i)
        continue
    if hdr and len(r) >= len(hdr):
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[c["# Samples"]] or 0), int(r[c["Instructions Executed"]] or 0),
                          int(r[c["L1 Wavefronts Shared"]] or 0)))
        except ValueError:
            pass
ts = sum(l[3] for l in lines); te = sum(l[4] for l in lines)
print("lines %d samples %d executed %d" % (len(lines), ts, te))
for l in sorted(lines, key=lambda x: -x[4])[:top]:
    print("%-16s %4d  exec%% %5.2f  samp%% %5.2f  wf %9d  %s" % (l[0], l[1], 100.0 * l[4] / te, 100.0 * l[3] / ts, l[5], l[2][:90]))
