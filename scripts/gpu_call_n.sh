#!/bin/bash
# multi-GPU call: NCCL parity test + bench at N GPUs (N = number of visible GPUs)
set -x
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q -s > gpurun_out/n${N}_nccl_test.log 2>&1; echo "rc=$?" >> gpurun_out/n${N}_nccl_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N $BENCH_ARGS > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?" >> gpurun_out/n${N}_bench.err
tail -3 gpurun_out/n${N}_nccl_test.log; tail -5 gpurun_out/n${N}_bench.err
