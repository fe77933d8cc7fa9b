"""Hot source lines of an `ncu --page source --csv --print-source=cuda,sass` export: share of stall samples and of
executed instructions per CUDA source line (all files of the kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and "# Samples" in hdr:
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[i_s] or 0), int(r[i_e] or 0)))
        except (ValueError, IndexError):
            pass
ts, te = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print("samples %d  warp-instructions %d" % (ts, te))
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%-18s %4d %5.1f%% smp %5.1f%% ins  %s" % (a[0], a[1], 100 * a[3] / ts, 100 * a[4] / te, a[2]))
