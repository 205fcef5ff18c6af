"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches.csv):
per-kernel totals / shares for ONE image and the launches in issue order.

    python profiles/summarize_launches.py gpurun_out/launches.csv [launches_per_image] > profiles/..._summary.txt

Only this repo's kernels (namespace sb::, shown by ncu as "unnamed>::k_*") are kept; the last
`launches_per_image` of them are one steady-state image (the bench runs the images back to back on one stream
with --contexts 1)."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)(<.*)?\(", name)
    if not m:
        return name[:60]
    base, targs = m.group(1), m.group(2) or ""
    if base == "k_stream":
        t = re.search(r"StreamGeom<([^>]*)>", targs)
        a = [x.strip().split(")")[-1] for x in t.group(1).split(",")] if t else []
        return f"k_stream<{','.join(a[1:4])}|C{a[4]} W{a[5]}>" if len(a) > 5 else "k_stream"
    if base in ("k_cascade", "k_blur", "k_extrema", "k_input_u8"):
        t = re.search(r"<([^>]*)>", targs)
        return f"{base}<{t.group(1).replace(' ', '')}>" if t else base
    return base


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        if "unnamed>::k_" not in r["Kernel Name"] and "sb::" not in r["Kernel Name"]:
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        rows.append((short(r["Kernel Name"]), r["Grid Size"], ns / 1e3))
    per_image = int(sys.argv[2]) if len(sys.argv) > 2 else None
    if per_image is None:
        # one image = from the last k_input_u8 launch to the end
        starts = [i for i, r in enumerate(rows) if r[0].startswith("k_input_u8")]
        per_image = len(rows) - starts[-1]
    img = rows[-per_image:]
    total = sum(r[2] for r in img)
    print(f"# {path}: {per_image} launches per image, sum of kernel times {total:.1f} us")
    print("# cold-cache, serialised per-launch times under ncu: compare SHARES, not absolutes")
    agg = OrderedDict()
    for name, _, us in img:
        base = name.split("<")[0] if not name.startswith(("k_cascade", "k_stream")) else name
        t, c = agg.get(base, (0.0, 0))
        agg[base] = (t + us, c + 1)
    print("  total us  count  share  kernel")
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {t:8.1f}  {c:5d}  {100 * t / total:4.1f}%  {name}")
    print("# per launch, in issue order")
    for name, grid, us in img:
        print(f"  {us:8.1f} us  {grid:<16s} {name}")


if __name__ == "__main__":
    main()
