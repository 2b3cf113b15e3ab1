"""Hot spots of one kernel from an ncu report's SASS source page: barriers with their wait samples, and every
instruction above a sample threshold.  python scripts/ncu_source_hot.py report.ncu-rep [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
data = rows[hi + 1:]
tot = sum(int(r[si]) for r in data)
print("instructions", len(data), "samples", tot)
run = 0
for i, r in enumerate(data):
    s = int(r[si])
    run += s
    if s >= thr or "BAR" in r[src]:
        print("%6d  samples %6d  cum %5.1f%%  exec %8s  %s" % (i, s, 100.0 * run / tot, r[ie], r[src].strip()[:80]))
