"""ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum --csv --log-file X`) -> markdown
table per kernel: launches, total ms, share, avg ms, grid, block.
usage: launch_list.py launches.csv out.md "title" """
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"\(.*$", "", name)
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache "
                "and serialised, compare SHARES.\n\n| kernel | launches | total ms | share | avg ms | grid | block |\n|---|---|---|---|---|---|---|\n")
        for name, (cnt, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:100]}` | {cnt} | {ms:.2f} | {100 * ms / total:.1f}% | {ms / cnt:.3f} | {grid} | {block} |\n")
        f.write(f"\ntotal {total:.1f} ms over {sum(a[0] for a in agg.values())} launches\n")


if __name__ == "__main__":
    main()
