"""Per-kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`).

    python scripts/launch_shares.py gpurun_out/launches.csv "title line" > profiles/rN_launch_shares.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(open(path, newline="")) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    k_i, m_i, u_i, v_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
    tot = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= v_i or r[m_i] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r[k_i])
        name = re.sub(r"\(.*$", "", name).replace("trajopt::", "").replace("(int)", "").replace("(bool)", "")
        ms = float(r[v_i].replace(",", "")) * scale[r[u_i]]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + ms)
    total = sum(t for _, t in tot.values())
    print(title)
    print()
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:60s} launches {n:5d}  total {t:10.3f} ms  share {100 * t / total:5.1f}%  avg {t / n:8.3f} ms")


if __name__ == "__main__":
    main()
