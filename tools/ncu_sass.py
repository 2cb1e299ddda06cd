"""Per-SASS-instruction execution share of one captured launch: python tools/ncu_sass.py x.ncu-rep [launch index]
(reads `ncu -i x --page source --csv`; needs a capture made with --import-source on / -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []
for r in rows:
    if r and r[0] == "Kernel Name": blocks.append([r]); continue
    if blocks: blocks[-1].append(r)
b = blocks[which]
hdr = b[1]; data = [r for r in b[2:] if len(r) == len(hdr)]
ia, it, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); tott = sum(int(r[it]) for r in data); smp = sum(int(r[ismp]) for r in data)
print(b[0][1][:100]); print("warp inst", tot, "thread inst", tott, "threads/inst %.2f" % (tott / tot), "sass", len(data), "samples", smp)
for i, r in enumerate(data):
    print("%4d %6.2f%% thr%5.1f smp%5.2f%%  %s" % (i, 100 * int(r[ia]) / tot, int(r[it]) / max(int(r[ia]), 1), 100 * int(r[ismp]) / max(smp, 1), r[isrc].strip()))
