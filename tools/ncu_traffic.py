#!/usr/bin/env python
"""Record the measured DRAM traffic (bytes per launch) of a kernel from an `ncu --set full`
report into profiles/traffic.json:   python tools/ncu_traffic.py <key> <file.ncu-rep>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(key, path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO("\n".join(l for l in out.splitlines() if not l.startswith("==")))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda n: float(vals[hdr.index(n)]) * UNIT[units[hdr.index(n)]]
    total = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    p = os.path.join(ROOT, "profiles", "traffic.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[key] = total
    json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print(key, total)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
