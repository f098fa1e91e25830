"""Static SASS statistics of one kernel in an object file, attributed to the kernel's stages (outermost inlined source
line): instruction count, LDL / STL (local-memory = spill or dynamically indexed array), FFMA share.
    python scripts/sass_static_stages.py build/obj/wb_inst_h1.o wb_linearize3 [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, mangled = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
for cub in sorted(os.listdir(td)):
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
    st = [i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." in l and mangled in l]
    if not st:
        continue
    stats = collections.defaultdict(lambda: collections.Counter())
    inner = collections.defaultdict(lambda: collections.Counter())
    chain, fresh = [], True
    for l in dis[st[0] + 1:]:
        if l.strip().startswith(".section") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(1).split(".")[0]
            k = chain[-1] if chain else ("?", 0)
            stats[k]["n"] += 1
            stats[k][op] += 1
            if op in ("LDL", "STL"):
                inner[chain[0] if chain else ("?", 0)][op] += 1
            fresh = True
    tot = sum(v["n"] for v in stats.values())
    print("kernel", mangled, "SASS", tot)
    for k, v in sorted(stats.items(), key=lambda x: -x[1]["n"])[:top]:
        print(f"  {k[0]}:{k[1]:<5d} sass {v['n']:6d}  FFMA {v['FFMA']:5d}  LDS {v['LDS']:5d}  LDG {v['LDG']:4d}  LDL {v['LDL']:4d}  STL {v['STL']:4d}")
    print("local-memory instructions by innermost line:")
    for k, v in sorted(inner.items(), key=lambda x: -(x[1]["LDL"] + x[1]["STL"]))[:12]:
        print(f"  {k[0]}:{k[1]:<5d} LDL {v['LDL']:4d} STL {v['STL']:4d}")
    break
