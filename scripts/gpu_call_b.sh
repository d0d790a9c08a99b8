#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
L=npr-sph_b200/lib
for v in "" _tpb32 _tpb128 _nomerge _tpb128nomerge; do
  NPRSPH_LIB=$PWD/$L/libnprsph$v.so timeout 300 python scripts/ab_profile.py 256 2000 >> gpurun_out/b_ab.jsonl 2>> gpurun_out/b_ab.err
done
for gap in 0.5 1.0 2.0; do timeout 300 python scripts/nan_probe.py 256 2500 250 $gap >> gpurun_out/b_nan.log 2>&1; done
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_rho$' --launch-skip 2000 -c 1 \
    -o gpurun_out/b_rho_evolved -f python scripts/profile_run.py 256 2000 2 > gpurun_out/b_ncu_rho.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_force_records$' --launch-skip 2000 -c 1 \
    -o gpurun_out/b_force_evolved -f python scripts/profile_run.py 256 2000 2 > gpurun_out/b_ncu_force.log 2>&1
tail -3 gpurun_out/b_tests.log
