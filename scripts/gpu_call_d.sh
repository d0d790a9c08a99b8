#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/d_tests.log
L=$PWD/npr-sph_b200/lib
for v in "" _flatwalk; do
  NPRSPH_LIB=$L/libnprsph$v.so timeout 300 python scripts/ab_profile.py 256 2000 >> gpurun_out/d_ab.jsonl 2>> gpurun_out/d_ab.err
done
timeout 900 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_bench.err
tail -3 gpurun_out/d_tests.log; cat gpurun_out/d_ab.jsonl
