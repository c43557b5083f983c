"""DRAM traffic per launch of every kernel in one or more .ncu-rep files (`ncu --set full` captures of the shipped build),
merged into profiles/ncu_traffic.json under "<kernel>@<sites>" -- the file bench.py reads for `roofline.traffic`.
    python tools/ncu_traffic.py <sites> gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(rep):
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    ki, ri, wi = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        base = re.sub(r"\(.*", "", r[ki]).split("::")[-1].strip()
        yield base, float(r[ri].replace(",", "")) * UNIT[units[ri]] + float(r[wi].replace(",", "")) * UNIT[units[wi]]


def main():
    n = int(sys.argv[1])
    acc = {}
    for rep in sys.argv[2:]:
        for k, b in launches(rep):
            acc.setdefault(k, []).append(b)
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for k, v in sorted(acc.items()):
        table[f"{k}@{n}"] = sum(v) / len(v)
        table[f"{k}@{n}:launches_captured"] = len(v)
        print(f"{k}@{n}: {sum(v) / len(v) / 1e6:.2f} MB per launch over {len(v)} captured launches")
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
