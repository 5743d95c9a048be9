#!/usr/bin/env python
"""Write profiles/traffic.json from ncu captures of the two sweep kernels (read here, no GPU needed):

    python profiles/ncu_traffic.py c3=<rep or csv> c2=<rep or csv> [note=...]

Each capture holds one launch of sweep_p16_kernel<..,COLS=1,..> and one of <..,COLS=0,..>
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:sweep_ -s N -c 2`); the value
stored per workload is read + write bytes of the pair = DRAM traffic of the sweep per VB iteration.
The file is stamped with the hash of the kernel sources (ccfindr_b200.build.kernel_hash): bench.py
refuses to print a traffic figure whose stamp does not match the sources it runs."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ccfindr_b200 import build as vb_build  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def rows_of(path):
    if path.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    else:
        out = open(path).read()
    lines = [ln for ln in out.splitlines() if ln.startswith('"')]
    return list(csv.reader(lines))


def traffic(path):
    rows = rows_of(path)
    hdr = rows[0]
    per_kernel = []
    if "Metric Name" in hdr:          # long format (--csv on the command line)
        ki, mi, ui, vi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
        ii = hdr.index("ID")
        acc = {}
        for r in rows[1:]:
            if r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                acc.setdefault((r[ii], r[ki]), 0.0)
                acc[(r[ii], r[ki])] += float(r[vi].replace(",", "")) * UNIT[r[ui]]
        per_kernel = [(k[1], v) for k, v in acc.items()]
    else:                              # raw page of a report: one row per launch, units in row 2
        units = rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(d[m].replace(",", "")) * UNIT[units[hdr.index(m)]]
            per_kernel.append((d["Kernel Name"], tot))
    sw = [(n, v) for n, v in per_kernel if "sweep_" in n]
    assert len(sw) >= 2, "need one launch of each sweep pass"
    return sum(v for _, v in sw[:2]), [n[:80] for n, _ in sw[:2]]


def main():
    out = {"source_hash": vb_build.kernel_hash()}
    notes = []
    for arg in sys.argv[1:]:
        k, v = arg.split("=", 1)
        if k == "note":
            notes.append(v)
            continue
        t, names = traffic(v)
        out[k] = int(t)
        notes.append("%s: %.3f GB over %s (%s)" % (k, t / 1e9, " + ".join(names), os.path.basename(v)))
    out["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the two sweep launches of one VB "
                   "iteration, ncu; " + "; ".join(notes))
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
