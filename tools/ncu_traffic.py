"""DRAM traffic per ray and kernel from a `ncu --set full` report of the bench (one launch per kernel on `rays` rays):
writes profiles/traffic_per_ray.json, which bench.py scales to the rays a launch of the timed run covers (`roofline.traffic`).

usage: python tools/ncu_traffic.py gpurun_out/<tag>_full.ncu-rep <rays per launch> <tag>
"""
import csv
import io
import json
import os
import subprocess
import sys

NAMES = [("knn_query_rays", "knn_query_rays"), ("visibility_kernel", "visibility"), ("aggregate_kernel", "aggregate"),
         ("fc_tail_kernel", "fc_tail"), ("neighbor2_kernel", "neighbor2"), ("row_gemm128_kernel<1>", "attn_tail"),
         ("row_gemm128_kernel<0>", "qproj"), ("ray2_kernel", "ray")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(path, rays, tag):
    # a .csv is the raw page already exported on the GPU box (`ncu -i rep --page raw --csv`): reports with many launches
    # are too large to bring back
    raw = open(path).read() if path.endswith(".csv") else subprocess.run(
        ["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    out, detail = {}, {}
    for r in rows[2:]:
        for pat, name in NAMES:
            if pat in r[ik]:                              # the last launch of a kernel wins (steady state, not setup)
                rd, wr = to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])
                out[name] = (rd + wr) / rays
                detail[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr}
    json.dump({"capture": tag, "rays_per_launch": rays, "bytes_per_ray": out, "per_launch": detail},
              open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic_per_ray.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3])
