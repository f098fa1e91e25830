"""Attribute the SASS-level samples / executed instructions of one kernel in an .ncu-rep to CUDA source lines using
nvdisasm --print-line-info on the cubin inside libmpxb.so (join by instruction order).
    python scripts/ncu_lines.py rep kernel_regex mangled_substring [top]"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, rx, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data0 = [r for r in rows if r and r[0].startswith("0x") and len(r) == len(hdr)]
seen = {}
data = []
for r in data0:  # the page lists every instruction once per collection pass: keep the first occurrence
    if r[0] not in seen:
        seen[r[0]] = 1
        data.append(r)
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("mpx_b200/libmpxb.so")], cwd=td, capture_output=True)
dis = []
for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):  # one cubin per translation unit
    d = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
    if any(".text." in l and mangled in l for l in d):
        dis = d
        break
# find the .text section of the kernel
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l or (l.strip().startswith(".section") and ".text." in l and mangled in l))
lines = []
cur = ("?", 0)
inl = None
for l in dis[start + 1:]:
    if l.strip().startswith(".section") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
print("sass in rep", len(data), "sass in cubin", len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for r, ln in zip(data, lines):
    agg[ln][0] += int(r[ix["# Samples"]] or 0)
    agg[ln][1] += int(r[ix["Instructions Executed"]] or 0)
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"{f:24s} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f}:{l:<5d} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%")
if os.environ.get("RANGES"):
    # RANGES="name:file:lo-hi,..." buckets
    bk = collections.defaultdict(lambda: [0, 0])
    spec = [s.split(":") for s in os.environ["RANGES"].split(",")]
    for (f, l), v in agg.items():
        nm = "other:" + f
        for name, ff, rng in spec:
            lo, hi = map(int, rng.split("-"))
            if f == ff and lo <= l <= hi:
                nm = name
                break
        bk[nm][0] += v[0]; bk[nm][1] += v[1]
    for k, v in sorted(bk.items(), key=lambda x: -x[1][0]):
        print(f"BUCKET {k:28s} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%")
