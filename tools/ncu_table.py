"""Print ncu --csv metric logs (long format) as one compact table per file.  usage: ncu_table.py file.csv..."""
import csv, sys
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if not rows:
        print(path, "empty"); continue
    hdr = rows[0]; iN, iV, iI, iK = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Kernel Name")
    d, names = {}, {}
    for r in rows[1:]:
        d.setdefault(r[iI], {})[r[iN]] = r[iV]; names[r[iI]] = r[iK].split("(")[0]
    ids = list(d)
    print("==", path)
    print("%-70s" % "metric", *["%14s" % (names[i][-12:] + "#" + i) for i in ids])
    for m in sorted({k for v in d.values() for k in v}):
        short = m.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")
        print("%-70s" % short[:70], *["%14s" % d[i].get(m, "-")[:14] for i in ids])
