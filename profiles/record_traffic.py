"""Turn an `ncu --set full --page raw --csv` export into an entry of profiles/r02_traffic.json:
DRAM bytes (read + write) of ONE launch of a kernel, the file it came from and the sha256 of the kernel's
source at capture time.  bench.py reports `roofline.traffic` from this file and refuses a stale capture.

    python profiles/record_traffic.py KEY profiles/<raw.csv> cli-p_b200/clipb200/csrc/<file>.cu [kernel-substring] [launch-index]
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(str(v).replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    key, raw, src = sys.argv[1:4]
    want = sys.argv[4] if len(sys.argv) > 4 else ""
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    rows = list(csv.reader(open(os.path.join(ROOT, raw))))
    header = next(r for r in rows if "Kernel Name" in r)
    units = rows[rows.index(header) + 1]
    col = {n: i for i, n in enumerate(header)}
    launches = [r for r in rows[rows.index(header) + 2:] if len(r) == len(header) and want in r[col["Kernel Name"]]]
    r = launches[which]
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    dur = r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]
    sha = hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest()[:16]
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    db[key] = {"bytes": rd + wr, "dram_read": rd, "dram_write": wr, "file": raw, "kernel": r[col["Kernel Name"]],
               "launch": f"launch {which} of {len(launches)} matching '{want}', grid {r[col.get('Grid Size', 0)]}, {dur}",
               "kernel_sha": sha, "kernel_source": src}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
    print(key, db[key])


if __name__ == "__main__":
    main()
