"""Aggregate the SASS source page of one kernel in an .ncu-rep: stall-reason shares, executed instructions by opcode
class, local-memory sectors.   python scripts/ncu_stalls.py rep kernel_regex"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows:
    if r and r[0].startswith("0x") and len(r) == len(hdr) and r[0] not in seen:
        seen.add(r[0]); data.append(r)
tot = {h: sum(int(r[ix[h]] or 0) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
s = sum(tot.values()) or 1
print("kernel", rx, "SASS instructions", len(data))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {k:28s} {100 * v / s:5.1f}%")
ops = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]] or 0)
ti = sum(ops.values())
print("  warp instructions executed", ti)
for k, v in ops.most_common(16):
    print(f"  {k:10s} {100 * v / ti:5.1f}%")
if "L2 Theoretical Sectors Local" in ix:
    print("  local-memory L2 sectors", sum(int(r[ix["L2 Theoretical Sectors Local"]] or 0) for r in data))
