"""Write profiles/traffic.json from an ncu --set full capture of `python bench.py` (the default workload):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the linearise and the LQ kernel.
    python scripts/ncu_traffic.py rep [workload]"""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "go2_wb"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    tot = sum(float(r[ix[k]]) * scale[units[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    key = "linearize" if "linearize" in name else ("lq_solve" if ("lq3" in name or "lq_seq" in name) else None)
    if key:
        out[key + "_dram_bytes_per_launch"] = tot
        out[key + "_kernel"] = name.split("(")[0].replace("void ", "")
out["capture"] = f"{os.path.basename(rep)} (ncu --set full --clock-control none, one launch per kernel inside `python bench.py`)"
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
table = json.load(open(path)) if os.path.exists(path) else {}
table[workload] = out
json.dump(table, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
