"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`)
as per-kernel time shares:  python tools/launch_shares.py X.csv [skip_launches]
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
            name = re.sub(r"^void\s+", "", r["Kernel Name"]).replace("<unnamed>::", "")
            name = re.sub(r"[<(].*", "", name)  # drop template / parameter lists
            rows.append((name, us))
    rows = rows[skip:]
    tot = sum(u for _, u in rows)
    agg, cnt = defaultdict(float), defaultdict(int)
    for k, u in rows:
        agg[k] += u
        cnt[k] += 1
    print(f"{len(rows)} launches, {tot:.1f} us summed")
    for k, u in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"  {u:9.1f} us  {100 * u / tot:5.1f}%  {cnt[k]:3d}x  {k}")


if __name__ == "__main__":
    main()
