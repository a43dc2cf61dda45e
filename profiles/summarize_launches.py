"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel share table kept in profiles/README.md.

    python profiles/summarize_launches.py gpurun_out/c6_launches_decode.csv [more.csv ...]

Times under ncu are serialised and cold-cache: only the SHARES are comparable with bench.py's CUDA-event numbers.
"""
import csv
import re
import sys
from collections import OrderedDict


def summarize(path: str) -> str:
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"].replace(",", ""))
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r.get("Metric Unit", "ns"), 1e-3)
            rows.append((r["Kernel Name"], ns * scale))
    agg = OrderedDict()
    for name, us in rows:
        short = re.sub(r"^(void\s+)?(<unnamed>::|\(anonymous namespace\)::)", "", name)
        a = agg.setdefault(short, [0.0, 0])
        a[0] += us
        a[1] += 1
    total = sum(v[0] for v in agg.values())
    out = [f"{path}: total {total:.1f} us over {len(rows)} launches"]
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"  {us:9.1f} us {100 * us / total:5.1f}%  n={n:3d} avg={us / n:8.2f}  {name[:110]}")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summarize(p))
        print()
