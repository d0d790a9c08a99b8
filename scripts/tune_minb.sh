#!/bin/bash
# A/B of launch-bounds register caps for the two neighbour passes (run on the GPU box)
cd "$(dirname "$0")/.."
for cfg in "8 4" "8 5" "8 8" "10 6" "6 6" "12 6"; do
  set -- $cfg
  touch npr-sph_b200/csrc/sph_passes.cu
  make -C npr-sph_b200 EXTRA_NVFLAGS="-DNPRSPH_RHO_MINB=$1 -DNPRSPH_FORCE_MINB=$2" > /dev/null 2>&1
  regs=$(grep -E "registers" npr-sph_b200/build/sph_passes.ptxas.log | sed -n '7p;9p' | sed -E 's/.*Used ([0-9]+) registers.*/\1/' | tr '\n' ' ')
  echo "RHO_MINB=$1 FORCE_MINB=$2 regs(rho_mask,force_mask)=$regs $(QP_ONLY=2,0 python scripts/quick_profile.py 256 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rho %.3f force %.3f step %.3f' % (d['stages_ms']['rho'], d['stages_ms']['force'], d['ms_per_step_wall']))")"
done
