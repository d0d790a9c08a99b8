#!/bin/bash
# A/B of compile-time variants of the neighbour passes on the GPU box:
#   scripts/ab.sh "-DNPRSPH_WALK_PIPE=0" "-DNPRSPH_WALK_PIPE=1" ...
# rebuilds libnprsph.so with each flag set and prints the per-stage times of the 16 Mi dam break.
cd "$(dirname "$0")/.."
for flags in "$@"; do
  touch npr-sph_b200/csrc/sph_passes.cu
  make -C npr-sph_b200 EXTRA_NVFLAGS="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "[$flags] $(QP_ONLY=2,0 python scripts/quick_profile.py ${AB_SIDE:-256} | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rho %.3f force %.3f step %.3f' % (d['stages_ms']['rho'], d['stages_ms']['force'], d['ms_per_step_wall']))")"
done
touch npr-sph_b200/csrc/sph_passes.cu
make -C npr-sph_b200 > /dev/null 2>&1
