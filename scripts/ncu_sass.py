"""Per-SASS-instruction executed counts (per warp) and stall samples of one kernel in an .ncu-rep."""
import csv, subprocess, sys, io
rep, kern, warps = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iA, iI, iS, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
lines = []
for r in rows[2:]:
    try: n = int(r[iI])
    except Exception: continue
    lines.append((r[iA].strip(), n, int(r[iS] or 0), r[iT]))
half = [k for k, l in enumerate(lines) if "EXIT" in l[0]]
print("total per warp", sum(l[1] for l in lines) / warps, "samples", sum(l[2] for l in lines))
for k, (s, n, sm, t) in enumerate(lines):
    if n / warps > thr:
        print(f"{k:4d} {n / warps:8.1f} {sm:6d} {t:>5s}  {' '.join(s.split()[:6])}")
