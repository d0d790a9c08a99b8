#!/bin/bash
# A/B of compile-time variants on the lattice AND the evolved dam break (2 Mi particles):
#   scripts/ab2.sh "-DNPRSPH_PAIR_DZ=1" "-DNPRSPH_PAIR_DZ=3" ...
cd "$(dirname "$0")/.."
for flags in "$@"; do
  touch npr-sph_b200/csrc/sph_passes.cu
  make -C npr-sph_b200 EXTRA_NVFLAGS="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "[$flags]"; python scripts/evolved_profile.py ${AB_SIDE:-128} 0 4000 | grep steps | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   step %5d: rho %.3f force %.3f step %.3f ms, pair walks %.4f' % (d['steps'], d['ms']['rho'], d['ms']['force'], d['step_ms'], d['pair_walk_frac']))"
done
touch npr-sph_b200/csrc/sph_passes.cu
make -C npr-sph_b200 > /dev/null 2>&1
