#!/bin/bash
# A/B of launch-bounds register caps for the two neighbour passes (run on the GPU box)
cd "$(dirname "$0")/.."
for cfg in ${TUNE_CFGS:-8,6 8,7 8,8 8,9}; do
  r=${cfg%,*}; f=${cfg#*,}
  touch npr-sph_b200/csrc/sph_passes.cu
  make -C npr-sph_b200 EXTRA_NVFLAGS="-DNPRSPH_RHO_MINB=$r -DNPRSPH_FORCE_MINB=$f" > /dev/null 2>&1
  echo "RHO_MINB=$r FORCE_MINB=$f $(QP_ONLY=2,0 python scripts/quick_profile.py 256 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rho %.3f force %.3f step %.3f' % (d['stages_ms']['rho'], d['stages_ms']['force'], d['ms_per_step_wall']))")"
done
