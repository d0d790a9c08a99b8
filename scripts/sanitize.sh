#!/bin/bash
# compute-sanitizer (memcheck, then racecheck + initcheck on a smaller set) over the GPU tests that run
# small scenes; the 1 M / 16 M tests are left out (the tool slows kernels 10-100x).
set -x
mkdir -p gpurun_out
SEL='not million and not sixteen and not long_axis and not nan_onset'
timeout 2400 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_consumer.py \
    -m gpu -q -x -k "$SEL" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_sort.py tests/test_gpu_api.py -m gpu -q -x -k "sort or fused or streaming or dense" \
    > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/san_racecheck.log
tail -5 gpurun_out/san_memcheck.log; tail -5 gpurun_out/san_racecheck.log
