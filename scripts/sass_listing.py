"""SASS listings of the hot Go2 / H1 / SRBD kernels in mpx_b200/libmpxb.so (cuobjdump -sass), encodings stripped:
    python scripts/sass_listing.py <tag> [name-prefix,...]
writes profiles/sass/<tag>_<kernel>.sass.gz (instruction text, one per line) and profiles/<tag>_sass_opcodes.txt
(instruction count and opcode histogram per kernel -- FFMA share, LDS/LDG/LDL counts, no HMMA/UTCMMA expected: the
path is FP32 CUDA-core work)."""
import collections, gzip, os, re, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mpx_b200", "libmpxb.so")
kernels = {
    "lq_seq_61_12": "_ZN4mpxb13lq_seq_kernelILi61ELi12ELi256EEEvNS_6LqArgsE",
    "lq_seq_75_19": "_ZN4mpxb13lq_seq_kernelILi75ELi19ELi256EEEvNS_6LqArgsE",
    "lq_seq_13_12": "_ZN4mpxb13lq_seq_kernelILi13ELi12ELi256EEEvNS_6LqArgsE",
    "lq_scan_61_12": "_ZN4mpxb14lq_scan_kernelILi61ELi12ELi256EEEvNS_6LqArgsE",
    "lq3_go2": "_ZN4mpxb10lq3_kernelILi37ELi24ELi12ELi4ELi128EEEvNS_7Lq3ArgsE",
    "lq3_h1": "_ZN4mpxb10lq3_kernelILi51ELi24ELi19ELi0ELi192EEEvNS_7Lq3ArgsE",
    "lq_small_13_12": "_ZN4mpxb15lq_small_kernelILi13ELi12ELi256EEEvNS_6LqArgsE",
    "wb_linearize3_quad": "_ZN4mpxb20wb_linearize3_kernelILi14ELi18ELi4ELi192ELi64ELi2EEEvPKNS_7WbModelENS_7ModelIOENS_6LinOutE",
    "wb_linearize3_h1": "_ZN4mpxb20wb_linearize3_kernelILi21ELi25ELi4ELi320ELi128ELi1EEEvPKNS_7WbModelENS_7ModelIOENS_6LinOutE",
    "wb_trial3_quad": "_ZN4mpxb16wb_trial3_kernelILi14ELi18ELi4ELi32EEEvPKNS_7WbModelENS_7ModelIOENS_8TrialOutE",
    "wb_trial3_h1": "_ZN4mpxb16wb_trial3_kernelILi21ELi25ELi4ELi64EEEvPKNS_7WbModelENS_7ModelIOENS_8TrialOutE",
    "srbd_linearize": "_ZN4mpxb21srbd_linearize_kernelENS_10SrbdParamsENS_7ModelIOENS_6LinOutE",
    "select_update": "_ZN4mpxb20select_update_kernelENS_10SelectArgsE",
    "pscan_combine_81": "_ZN4mpxb20pscan_combine_kernelILi81ELi512EEEviiPKiNS_8PscanBufES3_",
    "pscan_combine_13": "_ZN4mpxb20pscan_combine_kernelILi13ELi512EEEviiPKiNS_8PscanBufES3_",
    "pscan_init_81_46": "_ZN4mpxb17pscan_init_kernelILi81ELi46ELi256EEEvNS_6LqArgsENS_8PscanBufE",
    "pscan_gains_81_46": "_ZN4mpxb18pscan_gains_kernelILi81ELi46ELi256EEEvNS_6LqArgsENS_8PscanBufE",
}
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None   # optional: comma-separated name prefixes
if only:
    kernels = {k: v for k, v in kernels.items() if any(k.startswith(o) for o in only)}
os.makedirs(os.path.join(root, "profiles", "sass"), exist_ok=True)
rep = []
for name, sym in kernels.items():
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", sym, so], capture_output=True, text=True).stdout
    ins = []
    for l in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?;)", l)
        if m:
            ins.append(f"/*{m.group(1)}*/ {m.group(2).strip()}")
    if not ins:
        rep.append(f"== {name}: not found in {so}")
        continue
    with gzip.open(os.path.join(root, "profiles", "sass", f"{tag}_{name}.sass.gz"), "wt") as f:
        f.write(f"// {sym}\n" + "\n".join(ins) + "\n")
    ops = collections.Counter()
    for i in ins:
        body = i.split("*/", 1)[1].strip()
        body = re.sub(r"^@!?U?P(T|\d+)\s+", "", body)
        ops[body.split()[0].rstrip(";").split(".")[0]] += 1
    rep.append(f"== {name}  ({sym})\n   {len(ins)} SASS instructions (static)")
    rep.append("   " + "  ".join(f"{o} {c}" for o, c in ops.most_common(24)))
    tc = sum(c for o, c in ops.items() if o in ("HMMA", "IMMA", "UTCMMA", "UTCHMMA", "QGMMA", "HGMMA"))
    rep.append(f"   tensor-core instructions: {tc};  LDL {ops.get('LDL', 0)}  STL {ops.get('STL', 0)}  LDGSTS {ops.get('LDGSTS', 0)}  BAR {ops.get('BAR', 0)}")
open(os.path.join(root, "profiles", f"{tag}_sass_opcodes.txt"), "w").write("\n".join(rep) + "\n")
print("\n".join(rep))
