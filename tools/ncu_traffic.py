#!/usr/bin/env python
"""profiles/r2_spmv_dram_traffic.json from `ncu --set full` captures of one spmv launch per mode (read here, no GPU):
dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, tagged with the hash of the kernel sources the
captured library was built from -- bench.py reports `roofline.traffic` only while that hash equals the tree's.

    python tools/ncu_traffic.py secded=gpurun_out/x.ncu-rep crc32c=gpurun_out/y.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def main():
    out = {"source_hash": bench.source_hash(), "workload": "poisson3d_256", "what": "one spmv launch, ncu --set full --clock-control none"}
    for arg in sys.argv[1:]:
        mode, rep = arg.split("=", 1)
        text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        out[mode] = {"traffic": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": r[hdr.index("Kernel Name")],
                     "duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")), "capture": os.path.basename(rep)}
    with open(os.path.join(ROOT, "profiles", "r2_spmv_dram_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
