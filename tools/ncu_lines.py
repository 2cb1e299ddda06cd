"""Hottest CUDA source lines of one captured launch: python tools/ncu_lines.py x.ncu-rep [launch index] [top N]
(reads `ncu -i x --page source --print-source cuda --csv`; capture with --import-source on, build with -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# blocks: "Kernel Name" row starts a launch; inside, "File Name" rows start a file, followed by a header row
launches = []
for r in rows:
    if r and r[0] == "Kernel Name": launches.append([]); continue
    if launches: launches[-1].append(r)
L = launches[which] if launches else rows
items = []; fname = None; hdr = None
for r in L:
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: n = int(d.get("Instructions Executed", "0") or 0)
        except ValueError: continue
        if n: items.append((n, int(d.get("Thread Instructions Executed", "0") or 0), int(d.get("# Samples", "0") or 0), fname, d["Line No"], d["Source"].strip()[:110]))
tot = sum(i[0] for i in items); smp = sum(i[2] for i in items)
print("warp inst", tot, "samples", smp)
for n, t, s, f, ln, src in sorted(items, reverse=True)[:top]:
    print("%5.1f%% thr%5.1f smp%5.1f%%  %s:%s  %s" % (100 * n / tot, t / n, 100 * s / max(smp, 1), f, ln, src))
