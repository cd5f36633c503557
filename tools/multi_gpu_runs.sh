#!/bin/bash
# Multi-GPU measurement recipes (run under `gpurun --gpus N -- bash tools/multi_gpu_runs.sh <what>`); results go to gpurun_out/.
# One process per GPU through torch.distributed.run on 127.0.0.1.
set -u
what=${1:-n2}
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
mkdir -p gpurun_out
case $what in
  n2)   # 2 GPUs: NCCL == single GPU test, weak-scaling bench, CFG-split bench (= config 2 on two GPUs)
    timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
    run 2 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 1500 gpurun_out/r02_bench_n2.json
    run 2 bench.py --gpus 2 --steps 3 --warmup 2 --cfg-split > gpurun_out/r02_bench_n2_cfgsplit.json 2> gpurun_out/r02_bench_n2_cfgsplit.err; tail -c 1500 gpurun_out/r02_bench_n2_cfgsplit.json
    ;;
  c3)   # config 3 on 4 GPUs (48 frames, 4 windows, musev_referencenet + ReferenceNet one-shot + IP tokens)
    run 4 tools/run_config.py --config 3 2>&1 | grep RUN_CONFIG | tee gpurun_out/r02_config3_n4.json
    run 4 tools/run_config.py --config 3 --cfg-split 2>&1 | grep RUN_CONFIG | tee gpurun_out/r02_config3_n4_cfgsplit.json
    ;;
  c45)  # configs 4 and 5 on 8 GPUs
    run 8 tools/run_config.py --config 4 2>&1 | grep RUN_CONFIG | tee gpurun_out/r02_config4_n8.json
    run 8 tools/run_config.py --config 5 2>&1 | grep RUN_CONFIG | tee gpurun_out/r02_config5_n8.json
    run 8 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 1200 gpurun_out/r02_bench_n8.json
    ;;
  c1)   # single-GPU runs of the configs (for the balance / efficiency tables)
    python tools/run_config.py --config 3 2>&1 | grep RUN_CONFIG | tee gpurun_out/r02_config3_n1.json
    python tools/run_config.py --config 4 --frames 40 2>&1 | grep RUN_CONFIG | tee gpurun_out/r02_config4_n1_40frames.json
    ;;
esac
