"""Per-SASS-instruction execution counts of one captured launch: python tools/ncu_sass.py x.ncu-rep [launch index] > out.txt
Columns: offset, warp-level executions (millions), average active threads, stall samples, instruction."""
import csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []
for r in rows:
    if r and r[0] == "Kernel Name": blocks.append([r]); continue
    if blocks: blocks[-1].append(r)
b = blocks[which]; hdr = b[1]; data = [r for r in b[2:] if len(r) == len(hdr)]
ia, it, ismp, iad, isrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Address"), hdr.index("Source")
base = int(data[0][iad], 16)
tot = sum(int(r[ia]) for r in data); stot = sum(int(r[ismp]) for r in data)
print(b[0][1][:100], "warp inst %.3f G" % (tot / 1e9), "samples", stot)
for r in data:
    n = int(r[ia])
    print("%05x %9.2f %5.1f %6d  %s" % (int(r[iad], 16) - base, n / 1e6, int(r[it]) / max(n, 1), int(r[ismp]), r[isrc].strip()))
