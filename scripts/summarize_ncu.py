"""Summaries of ncu CSV exports for profiles/.

    python scripts/summarize_ncu.py launches <launch-list.csv>      per-kernel launch count / total ms / share (markdown)
    python scripts/summarize_ncu.py raw <raw-page.csv> [regex]      selected metrics of each captured launch (markdown)
"""
import csv
import io
import re
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
]


def _rows(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    return list(csv.reader(io.StringIO(text[start:])))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?!bool|int).*$", "", name)
    return name.replace("bocf::", "").replace("sg::", "")


def launches(path):
    rows = _rows(path)
    hdr = rows[0]
    k, mname, val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= val or r[mname] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r[k]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[val].replace(",", "")) / 1e6
    ours = {n: a for n, a in agg.items() if not n.startswith(("cutlass", "at::", "std::", "vectorized", "elementwise")) and "cublas" not in n
            and "at::native" not in n}
    tot = sum(a[1] for a in ours.values())
    print("| kernel | launches | total ms | share of our kernels |\n|---|---:|---:|---:|")
    for n, a in sorted(ours.items(), key=lambda t: -t[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (n, a[0], a[1], 100 * a[1] / tot))
    print("\nTotal %.1f ms over %d launches of our kernels; other (library / torch) launches in the list: %d." % (
        tot, sum(a[0] for a in ours.values()), sum(a[0] for n, a in agg.items() if n not in ours)))


def raw(path, pattern=None):
    rows = _rows(path)
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in KEEP if m in hdr]
    sel = [r for r in rows[2:] if len(r) > kcol and (pattern is None or re.search(pattern, r[kcol]))]
    names = ["%s #%s" % (short(r[kcol]), r[0]) for r in sel]
    print("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---:|" * len(sel))
    for m, c in cols:
        print("| %s | %s | " % (m, units[c]) + " | ".join(r[c] for r in sel) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
