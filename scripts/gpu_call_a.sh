#!/bin/bash
# First GPU call of round 2: tests, smoke, bench, CUB yard-stick, launch list, ncu captures (1 GPU).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 120 scripts/cub_sort_yardstick.bin > gpurun_out/a_cub.json 2>&1
# launch list of a short evolved run (shares, cold cache): 3 steps after 400 evolve steps at 16 Mi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 60 --csv \
    --log-file gpurun_out/a_launches.csv python scripts/profile_run.py 256 500 4 > gpurun_out/a_launches.log 2>&1
# the two neighbour kernels on the evolved fluid (step 2000), full counter set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rho<' --launch-skip 2000 -c 1 \
    -o gpurun_out/a_rho_evolved -f python scripts/profile_run.py 256 2000 2 > gpurun_out/a_ncu_rho.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force_records' --launch-skip 2000 -c 1 \
    -o gpurun_out/a_force_evolved -f python scripts/profile_run.py 256 2000 2 > gpurun_out/a_ncu_force.log 2>&1
tail -3 gpurun_out/a_tests.log
