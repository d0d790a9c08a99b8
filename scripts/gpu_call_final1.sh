#!/bin/bash
# single-GPU evidence for the round: tests, smoke, bench, launch list, ncu captures of the evolved fluid
set -x
mkdir -p gpurun_out
P=gpurun_out/z
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "free_running" > ${P}_drift.log 2>&1
timeout 200 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 40 --csv \
    --log-file ${P}_launches.csv python scripts/profile_run.py 256 2000 6 > ${P}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_rho$' --launch-skip 2000 -c 1 \
    -o ${P}_rho_evolved -f python scripts/profile_run.py 256 2000 2 > ${P}_ncu_rho.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_force_records$' --launch-skip 2000 -c 1 \
    -o ${P}_force_evolved -f python scripts/profile_run.py 256 2000 2 > ${P}_ncu_force.log 2>&1
timeout 900 ncu --set full --clock-control none -k 'regex:^(k_rho|k_force_records|k_onesweep|k_reorder_cells|k_radix_hist|k_fill_gaps)$' --launch-skip 40 -c 8 \
    -o ${P}_step_lattice -f python scripts/profile_run.py 256 5 2 > ${P}_ncu_lattice.log 2>&1
tail -3 ${P}_tests.log
