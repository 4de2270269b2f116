"""profiles/r02_ncu_traffic.json from `ncu -i X.ncu-rep --page raw --csv` exports of the `--set full` captures: DRAM bytes
(read + write) per launch, per kernel family.  bench.py reads `gemm_dram_bytes_per_launch` for roofline.traffic.
    python scripts/make_ncu_traffic.py "cfg-2 ModCR PMR training shape" gpurun_out/prof_gemm.csv [more.csv ...]"""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload, paths = sys.argv[1], sys.argv[2:]
fam = {}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for path in paths:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])
        name = re.sub(r"^void ", "", name).split("<")[0].replace("modcr::", "")
        rd = float(r[ci["dram__bytes_read.sum"]].replace(",", "") or 0) * UNIT.get(units[ci["dram__bytes_read.sum"]], 1.0)
        wr = float(r[ci["dram__bytes_write.sum"]].replace(",", "") or 0) * UNIT.get(units[ci["dram__bytes_write.sum"]], 1.0)
        dur = float(r[ci["gpu__time_duration.sum"]].replace(",", "") or 0)
        f = fam.setdefault(name, dict(launches=0, read=0.0, write=0.0, dur=0.0))
        f["launches"] += 1; f["read"] += rd; f["write"] += wr; f["dur"] += dur
out_path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
ent = data.setdefault(workload, {})
for name, f in fam.items():
    ent[name] = dict(launches=f["launches"], dram_read_bytes_per_launch=f["read"] / f["launches"],
                     dram_write_bytes_per_launch=f["write"] / f["launches"])
if "gemm_kernel" in fam:
    g = fam["gemm_kernel"]
    ent["gemm_launches"] = g["launches"]
    ent["gemm_dram_bytes_per_launch"] = (g["read"] + g["write"]) / g["launches"]
json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(ent, indent=1))
