"""Attribute samples / executed instructions of one kernel in an .ncu-rep to the OUTERMOST source line (the line of
the __global__ function that the instruction was inlined into), i.e. to the kernel's stages.
    python scripts/ncu_stages.py rep kernel_regex mangled_substring [top]"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, rx, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows:
    if r and r[0].startswith("0x") and len(r) == len(hdr) and r[0] not in seen:
        seen.add(r[0]); data.append(r)
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("mpx_b200/libmpxb.so")], cwd=td, capture_output=True)
lines = None
for cub in sorted(os.listdir(td)):
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
    st = [i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." in l and mangled in l]
    if not st:
        continue
    lines, chain = [], []
    fresh = True
    for l in dis[st[0] + 1:]:
        if l.strip().startswith(".section") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines.append(tuple(chain) if chain else (("?", 0),))
            fresh = True
    break
print("sass in rep", len(data), "sass in cubin", len(lines))
outer = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r, ch in zip(data, lines):
    k = ch[-1]
    outer[k][0] += int(r[ix["# Samples"]] or 0)
    outer[k][1] += int(r[ix["Instructions Executed"]] or 0)
    outer[k][2] += int(r[ix["Thread Instructions Executed"]] or 0) if "Thread Instructions Executed" in ix else 0
    outer[k][3] += 1  # static SASS instructions of the stage (code size: 16 bytes each)
    outer[k][4] += int(r[ix["stall_no_inst"]] or 0) if "stall_no_inst" in ix else 0
ts = sum(v[0] for v in outer.values()) or 1
ti = sum(v[1] for v in outer.values()) or 1
print("total warp instructions", ti, "samples", ts)
if os.environ.get("INNER"):
    # INNER=<outer line>: break that stage down by the next frame of the inline chain
    tgt = int(os.environ["INNER"])
    inner = collections.defaultdict(lambda: [0, 0])
    for r, ch in zip(data, lines):
        if ch[-1][1] != tgt:
            continue
        k = ch[-2] if len(ch) > 1 else ch[-1]
        inner[k][0] += int(r[ix["# Samples"]] or 0)
        inner[k][1] += int(r[ix["Instructions Executed"]] or 0)
    for k, v in sorted(inner.items(), key=lambda x: -x[1][0])[:top]:
        print(f"  inner {k[0]}:{k[1]:<5d} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%")
    sys.exit(0)
for k, v in sorted(outer.items(), key=lambda x: -x[1][0])[:top]:
    lanes = v[2] / v[1] if v[1] else 0
    print(f"{k[0]}:{k[1]:<5d} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%  lanes/inst {lanes:4.1f}  sass {v[3]:6d}  no_inst {100 * v[4] / max(v[0], 1):4.1f}% of its samples")
