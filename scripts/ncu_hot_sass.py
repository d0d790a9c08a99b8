"""Per-SASS-instruction view of an .ncu-rep (read here, no GPU): executed warp instructions, lanes
active and stall samples, grouped into the straight-line blocks of the kernel so the hot loops show
up.   python scripts/ncu_hot_sass.py REPORT [min_share_percent]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > ix["Instructions Executed"]]
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(f"# {rows[0][1][:90]}")
print(f"# warp instructions {tot_i:,}  stall samples {tot_s:,}")
# blocks: split at branch instructions and at branch targets is overkill; split at BRA / BSYNC
blk, blocks = [], []
for r in body:
    blk.append(r)
    op = r[ix["Source"]].split()
    if any(t in ("BRA", "BSYNC.RECONVERGENT", "BSYNC", "EXIT") or t.startswith("BRA") for t in op[:2]):
        blocks.append(blk); blk = []
if blk: blocks.append(blk)
for b in blocks:
    ni = sum(int(r[ix["Instructions Executed"]] or 0) for r in b)
    ns = sum(int(r[ix["# Samples"]] or 0) for r in b)
    if 100.0 * ni / tot_i < min_share and 100.0 * ns / max(tot_s, 1) < min_share:
        continue
    thr = sum(int(r[ix["Thread Instructions Executed"]] or 0) for r in b)
    first = b[0][ix["Address"]][-4:]
    print(f"\nblock @{first}: {len(b):3d} instr, {100.0 * ni / tot_i:5.1f} % of warp instructions, {100.0 * ns / max(tot_s,1):5.1f} % of stall samples, "
          f"{thr / max(ni, 1):4.1f} lanes active, executed {int(b[0][ix['Instructions Executed']] or 0):,} times")
    top = sorted(b, key=lambda r: -int(r[ix["# Samples"]] or 0))[:4]
    for r in top:
        print(f"      {int(r[ix['# Samples']] or 0):7d} samples  {r[ix['Source']].strip()[:90]}")
