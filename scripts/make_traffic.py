#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full ... --page raw --csv` export: DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of a kernel
family) for the kernels bench.py reports a roofline for. usage: make_traffic.py raw.csv out.json source-note [cells-per-launch]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
iname = hdr.index("Kernel Name")
ird, iwr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
igrid = hdr.index("launch__grid_size")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def family(name):
    if "ImplicitTmaKernel" in name:
        # template argument MODE: 0 = dplur, 1 = axmb (implicit_tma.cuh)
        return "matrix_residual" if name.rstrip(">)").split(",")[-1].strip().startswith("(int)1") or ", 1>" in name else "dplur_sweep"
    if "ResidualMarchKernel" in name: return "residual"
    if "UpdateKernel" in name: return "update_norms"
    if "LusgsPencilKernel" in name: return "lusgs_plane"
    return None
acc = {}
for r in rows[2:]:
    f = family(r[iname])
    if not f: continue
    b = float(r[ird]) * scale[units[ird]] + float(r[iwr]) * scale[units[iwr]]
    acc.setdefault(f, []).append(b)
cells = int(sys.argv[4]) if len(sys.argv) > 4 else 256 ** 3
out = {f: {"cells": cells, "dram_bytes_per_launch": int(sum(v) / len(v)), "launches_captured": len(v), "source": sys.argv[3]}
       for f, v in acc.items()}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
