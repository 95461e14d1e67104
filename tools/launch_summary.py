"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py list.csv out.md "title" """
import collections
import csv
import sys


def main(path, out, title):
    rows = list(csv.reader(open(path)))
    hdr = None
    per = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[hdr.index("Metric Name")] == "gpu__time_duration.sum":
            name = r[hdr.index("Kernel Name")]
            v = float(r[hdr.index("Metric Value")].replace(",", ""))
            unit = r[hdr.index("Metric Unit")]
            us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
            per.setdefault(name, []).append(us)
    total = sum(sum(v) for v in per.values())
    L = [f"# {title}", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none --csv` - per-launch times are cold-cache and serialised; "
         "compare shares.", "", "| kernel | launches | mean us | max us | total ms | share |", "|---|---|---|---|---|---|"]
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        L.append(f"| `{name[:110]}` | {len(v)} | {sum(v) / len(v):.1f} | {max(v):.1f} | {sum(v) / 1e3:.2f} | {100 * sum(v) / total:.1f}% |")
    open(out, "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
