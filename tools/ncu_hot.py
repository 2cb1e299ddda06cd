"""Hot CUDA source lines of one captured launch, by joining ncu's SASS page with nvdisasm's line table.

    cuobjdump -xelf all sailor_b200/libsailor_pt_cuda.so ; nvdisasm -gi -c capi.sm_100a.cubin > capi_gi.sass
    python tools/ncu_hot.py x.ncu-rep <launch index> capi_gi.sass <substring of the mangled kernel name> [top N]

Prints two rankings: by innermost source line, and by call-site chain (innermost <- ... <- kernel body)."""
import collections, csv, io, re, subprocess, sys
rep, which, sass, key = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
# ---- nvdisasm: instruction offset -> chain of (file, line)
chains = {}; cur = []; pending = []; inside = False
for ln in open(sass):
    if ln.startswith("\t.section\t.text.") or ln.startswith(".text."):
        inside = key in ln
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        if pending: cur = pending; pending = []
        chains[int(m.group(1), 16)] = cur
# ---- ncu SASS page
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []
for r in rows:
    if r and r[0] == "Kernel Name": blocks.append([r]); continue
    if blocks: blocks[-1].append(r)
b = blocks[which]; hdr = b[1]; data = [r for r in b[2:] if len(r) == len(hdr)]
ia, it, ismp, iad = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
base = int(data[0][iad], 16)
inner = collections.Counter(); chainc = collections.Counter(); smp_inner = collections.Counter(); thr_inner = collections.Counter()
tot = 0; stot = 0
for r in data:
    off = int(r[iad], 16) - base
    c = chains.get(off, [("?", 0)])
    n = int(r[ia]); s = int(r[ismp]); tot += n; stot += s
    k = "%s:%d" % c[0] if c else "?"
    inner[k] += n; smp_inner[k] += s; thr_inner[k] += int(r[it])
    chainc[" <- ".join("%s:%d" % x for x in c[:4])] += n
print(b[0][1][:120]); print("warp inst", tot, "samples", stot, "sass", len(data), "mapped", sum(1 for r in data if (int(r[iad], 16) - base) in chains))
print("---- by innermost line")
for k, n in inner.most_common(top): print("%5.1f%% thr%5.1f smp%5.1f%%  %s" % (100 * n / tot, thr_inner[k] / max(n, 1), 100 * smp_inner[k] / max(stot, 1), k))
print("---- by inline chain")
for k, n in chainc.most_common(top): print("%5.1f%%  %s" % (100 * n / tot, k))
