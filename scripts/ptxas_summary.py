"""Registers / spills per kernel from the -Xptxas -v logs of the last build (npr-sph_b200/build/*.ptxas.log)."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "npr-sph_b200", "build")
for path in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    if len(sys.argv) > 1 and sys.argv[1] not in path:
        continue
    log = open(path).read()
    for e in re.split(r"ptxas info    : Compiling entry function '", log)[1:]:
        name = e.split("'")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("nprsph::(anonymous namespace)::", "").replace("void ", "")
        dem = re.sub(r"\(.*", "", dem)
        m = re.search(r"Used (\d+) registers", e)
        sp = re.search(r"(\d+) bytes spill stores", e)
        sm = re.search(r"(\d+) bytes smem", e)
        print(f"{os.path.basename(path)[:-10]:12s} {dem:60s} regs={m.group(1):>3s} spill={sp.group(1) if sp else '?':>3s} smem={sm.group(1) if sm else 0}")
