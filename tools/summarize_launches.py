"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, last_fraction=0.5):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], us))
    n = len(rows)
    rows = rows[int(n * (1 - last_fraction)):]          # keep the last iteration(s) only
    tot = defaultdict(lambda: [0, 0.0])
    for _, name, us in rows:
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
        name = re.sub(r"\(.*$", "", name)[:90]
        tot[name][0] += 1
        tot[name][1] += us
    total = sum(v[1] for v in tot.values())
    print(f"launches analysed: {len(rows)} (last {last_fraction:.0%} of {n}); summed kernel time {total / 1e3:.2f} ms\n")
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{name}` | {c} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / c:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
