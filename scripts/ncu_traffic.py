#!/usr/bin/env python
"""ncu report -> profiles/traffic_<workload>_<n>_<prec>.json (what bench.py reports as roofline.traffic).

  python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep c2 1024 f64 step3 [--stamp-from DIR]

dram_bytes_per_launch = mean over the captured launches of dram__bytes_read.sum + dram__bytes_write.sum
(`ncu --set full`), csrc_stamp = hash of the kernel sources of the build the capture was taken on
(DIR: root of the source tree that was sent to the GPU box; default: this checkout)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stamp(root):
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(root, "meep_b200", "csrc")
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    with open(os.path.join(root, "include", "meep_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()[:12]


def main():
    rep, workload, n, prec, kernel = sys.argv[1:6]
    root = sys.argv[sys.argv.index("--stamp-from") + 1] if "--stamp-from" in sys.argv else ROOT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    launches = []
    for r in rows[2:]:
        def val(name, sc):
            i = hdr.index(name)
            return float(r[i]) * sc[units[i]]
        launches.append({"kernel": r[hdr.index("Kernel Name")],
                         "dram_bytes": val("dram__bytes_read.sum", scale) + val("dram__bytes_write.sum", scale),
                         "dram_bytes_read": val("dram__bytes_read.sum", scale),
                         "dram_bytes_write": val("dram__bytes_write.sum", scale),
                         "seconds_under_ncu": val("gpu__time_duration.sum", tscale),
                         "registers": int(r[hdr.index("launch__registers_per_thread")]),
                         "grid": int(r[hdr.index("launch__grid_size")])})
    how = ("ncu --clock-control none with --set full, or a metrics-only pass (dram__bytes_read.sum, dram__bytes_write.sum, "
           "gpu__time_duration.sum, launch__registers_per_thread, launch__grid_size); dram__bytes_read.sum + dram__bytes_write.sum per launch")
    per_launch = sum(l["dram_bytes"] for l in launches) / len(launches)
    lean = [i for i, l in enumerate(launches) if "step3_lean_kernel" in l["kernel"]]
    if lean:
        # the fast path of one time step = two plan runs (B half: lean + slab + column launches; D-E half:
        # masked march or the same three): take the launches from one lean kernel of the B half up to the
        # next B half and halve their sum, so that the figure compares with alg_bytes_per_launch
        end = next((i for i in lean[1:] if (i - lean[0]) >= 4), len(launches))
        step = launches[lean[0]:end]
        if len(lean) >= 3 and lean[1] - lean[0] == 3:  # every half-step lean (single precision): two lean kernels per step
            end = lean[2]
            step = launches[lean[0]:end]
        per_launch = sum(l["dram_bytes"] for l in step) / 2.0
        launches = step
        how += "; one time step of fast-path launches (%d kernels) summed and halved (two plan runs per step)" % len(step)
    doc = {"workload": workload, "n": int(n), "prec": prec, "kernel": kernel,
           "dram_bytes_per_launch": per_launch,
           "launches": launches, "csrc_stamp": stamp(root), "report": os.path.basename(rep),
           "how": how}
    path = os.path.join(ROOT, "profiles", "traffic_%s_%s_%s.json" % (workload, n, prec))
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=1)
    print(path, doc["dram_bytes_per_launch"], doc["csrc_stamp"])


if __name__ == "__main__":
    main()
