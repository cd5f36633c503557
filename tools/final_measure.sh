#!/bin/bash
# One-GPU evidence run of a round: GPU suite, smoke, both bench arms, forward split, ncu launch list of the bench command and
# `ncu --set full` captures of the dominant kernels. Everything lands in gpurun_out/ (copy what is to be judged into profiles/).
#   gpurun --timeout 1700 -- 'bash tools/final_measure.sh r02'
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
python -m pytest tests -q -m gpu 2>&1 | tail -25 > $o/${tag}_pytest_gpu_final.log
tail -3 $o/${tag}_pytest_gpu_final.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee $o/${tag}_smoke.log
python bench.py --impl reference 2> $o/${tag}_bench_reference.err | tail -1 > $o/${tag}_bench_reference.json
python bench.py 2> $o/${tag}_bench_final.err | tail -1 > $o/${tag}_bench_final.json
cat $o/${tag}_bench_final.json
python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2> /dev/null | tail -1 > $o/${tag}_bench_final_k20.json
for p in musev musev_referencenet; do python tools/gpu_time_forward.py --preset $p --iters 10 --tag final 2>&1 | grep FWD_TIME; done | tee $o/${tag}_forward_split.txt
MVB_TRACE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/launches.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > /dev/null 2> $o/${tag}_launches_trace.log
python tools/analyze_launches.py /tmp/launches.csv $o/${tag}_launches_trace.log $o/${tag}_launches_by_shape.txt | head -12
# the .ncu-rep stays on the box (gpurun_out/ is limited to 64 MiB and drops EVERYTHING beyond it): only its text summary returns
PROF_REPS=1 ncu --set full --clock-control none -o /tmp/kernels_full -f python tools/gpu_prof_kernels.py > /tmp/ncu_full.log 2>&1
tail -2 /tmp/ncu_full.log
python tools/ncu_summary.py /tmp/kernels_full.ncu-rep > $o/${tag}_ncu_kernels_full_summary.txt
