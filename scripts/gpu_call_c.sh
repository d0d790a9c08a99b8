#!/bin/bash
set -x
mkdir -p gpurun_out
L=$PWD/npr-sph_b200/lib
for v in "" _f32w _f24w _r36w _dz2 _dz5; do
  NPRSPH_LIB=$L/libnprsph$v.so timeout 300 python scripts/ab_profile.py 256 2000 >> gpurun_out/c_ab.jsonl 2>> gpurun_out/c_ab.err
done
timeout 600 python scripts/config_lines.py config1 h4s_16M > gpurun_out/c_config_lines.jsonl 2> gpurun_out/c_config_lines.err
tail -2 gpurun_out/c_config_lines.err
