#!/usr/bin/env python
"""Write profiles/gram_fused_traffic.json (what bench.py reports as roofline.traffic) from an `ncu --set full` capture:

    ncu -i gpurun_out/gram_fused_r02.ncu-rep --page raw --csv > gpurun_out/raw_r02.csv
    python tools/update_traffic.py gpurun_out/raw_r02.csv 1000000

The dominant kernel of the benchmarked configuration is gram_struct_kernel (G1-12dof is a legged tree); gram_fused_kernel is matched
for captures of the unstructured path.  The JSON carries a content stamp of the kernel sources (bench.KERNEL_SOURCES) and the git
sha of HEAD: bench.py reports the traffic only while the stamp matches the tree it runs from."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    path, samples = sys.argv[1], int(sys.argv[2])
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    kcol = names.index("Kernel Name")
    body = [r for r in rows[hdr + 2:] if len(r) == len(names) and ("gram_struct" in r[kcol] or "gram_fused" in r[kcol])]
    if not body:
        raise SystemExit("no gram_struct_kernel / gram_fused_kernel launch in " + path)
    r = body[-1]
    kernel = "gram_struct_kernel" if "gram_struct" in r[kcol] else "gram_fused_kernel"

    def metric(name):
        i = names.index(name)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    from bench import kernel_source_stamp
    sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    out = {"kernel": kernel, "samples_per_launch": samples,
           "dram_bytes_read": int(metric("dram__bytes_read.sum")), "dram_bytes_write": int(metric("dram__bytes_write.sum")),
           "gpu_time_ms": metric("gpu__time_duration.sum") / (1e6 if units[names.index("gpu__time_duration.sum")] in ("nsecond", "ns") else 1.0),
           "source_stamp": kernel_source_stamp(), "git_sha": sha,
           "source": f"ncu --set full --clock-control none capture ({os.path.basename(path)})"}
    with open(os.path.join(ROOT, "profiles", "gram_fused_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
