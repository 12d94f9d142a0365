"""Pretty-print an ncu --csv --metrics log (scripts/k1_metrics.sh): one row per launch, one column per metric."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
by = collections.OrderedDict()
for r in rows[1:]:
    by.setdefault((r[iid], r[ik][:70]), {})[r[im]] = r[iv]
short = lambda m: m.replace("smsp__average_warps_issue_stalled_", "st_").replace("_per_issue_active.ratio", "").replace(".sum", "").replace(".avg.pct_of_peak_sustained_active", "%").replace("l1tex__data_", "").replace("pipe_lsu_", "")
for (i, k), m in by.items():
    print(k)
    print("   " + "  ".join("%s=%s" % (short(a), b) for a, b in m.items()))
