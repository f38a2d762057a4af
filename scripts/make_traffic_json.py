#!/usr/bin/env python
"""profiles/dominant_kernel_traffic.json from an `ncu --set full` report: measured DRAM bytes (read + write) per launch,
summed over the kernels of each bench.py stage.  usage: python scripts/make_traffic_json.py gpurun_out/x.ncu-rep"""
import csv
import io
import json
import subprocess
import sys

STAGE_OF = {"k_prefilter8": "cost", "k_cost_fused": "cost", "k_prefilter_tab": "cost", "k_cost_tma": "cost", "k_vertical3": "vertical", "k_hfwd": "horizontal", "k_hrev": "horizontal",
            "k_points_fuse": "fuse", "k_select_fused": "select", "k_cc_apply_bands": "post"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    per_kernel = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        key = next((k for k in STAGE_OF if k in name), None)
        if key is None or key in per_kernel:
            continue   # first profiled launch of each kernel
        b = sum(float(r[idx[m]].replace(",", "")) * SCALE[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per_kernel[key] = b
    out = {}
    for k, b in per_kernel.items():
        out[STAGE_OF[k]] = out.get(STAGE_OF[k], 0.0) + b
    out = {k: round(v) for k, v in out.items()}
    out["_source"] = f"{rep}: dram__bytes_read.sum + dram__bytes_write.sum per launch, B = 33 frames per launch"
    out["_frames_per_launch"] = 33
    out["_kernels"] = {k: round(v) for k, v in per_kernel.items()}
    json.dump(out, open("profiles/dominant_kernel_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
