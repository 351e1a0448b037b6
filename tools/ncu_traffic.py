"""profiles/ncu_traffic.json from .ncu-rep files: DRAM bytes per k-point of the dominant kernels (bench.py `roofline.traffic`).

    python tools/ncu_traffic.py OUT.json  REP:workload:class:kpoints_per_launch:kernel-regex[:launches_per_chunk] ...

For every spec the launches in REP whose kernel name matches the regex are summed (dram__bytes_read.sum +
dram__bytes_write.sum of an `ncu --set full --clock-control none` capture) and divided by the k-points they processed:
`kpoints_per_launch` k-points per `launches_per_chunk` matching launches (a class made of several kernels per chunk, e.g.
the staged tridiagonalisation, counts all of them).  Run in the build container (needs ncu to read the report).
"""
import csv
import io
import json
import re
import subprocess
import sys


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    out_path = sys.argv[1]
    try:
        table = json.load(open(out_path))
    except Exception:  # noqa: BLE001
        table = {}
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        rep, workload, cls, kpl, rx = parts[:5]
        per_chunk = int(parts[5]) if len(parts) > 5 else 1
        hdr, units, rows = rows_of(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        tot = 0.0
        n = 0
        names = set()
        for r in rows:
            name = r[ix["Kernel Name"]]
            if not re.search(rx, name):
                continue
            names.add(name.split("(")[0][-70:])
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += to_bytes(r[ix[m]], units[ix[m]])
            n += 1
        if n == 0:
            print(f"{spec}: no matching launches", file=sys.stderr)
            continue
        chunks = n / per_chunk
        table[f"{workload}:{cls}"] = {
            "bytes_per_kpoint": tot / (chunks * float(kpl)),
            "source": f"{rep.split('/')[-1]} (ncu --set full --clock-control none; {n} launches of {sorted(names)}, "
                      f"{kpl} k-points per {per_chunk} launch(es); dram__bytes_read.sum + dram__bytes_write.sum)",
        }
        print(f"{workload}:{cls}: {table[f'{workload}:{cls}']['bytes_per_kpoint']:.1f} B per k-point from {n} launches")
    json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
