#!/bin/bash
# Second half of the evidence run (the first lost its files to the 64 MiB gpurun_out limit): reference arm, launch list by shape,
# ncu --set full captures summarised ON the box (only text comes back).
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
python bench.py --impl reference 2> $o/${tag}_bench_reference.err | tail -1 > $o/${tag}_bench_reference.json
cat $o/${tag}_bench_reference.json
MVB_TRACE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > /dev/null 2> /tmp/launches_trace.log
python tools/analyze_launches.py /tmp/launches.csv /tmp/launches_trace.log $o/${tag}_launches_by_shape.txt | head -5
head -3 /tmp/launches.csv > $o/${tag}_launches_head.csv
PROF_REPS=1 ncu --set full --clock-control none -o /tmp/kernels_full -f python tools/gpu_prof_kernels.py > /tmp/ncu_full.log 2>&1
tail -2 /tmp/ncu_full.log
python tools/ncu_summary.py /tmp/kernels_full.ncu-rep > $o/${tag}_ncu_kernels_full_summary.txt
ls -la /tmp/kernels_full.ncu-rep
grep -c "^==" $o/${tag}_ncu_kernels_full_summary.txt
./tools/microbench/fmnmx_rate | tee $o/${tag}_fmnmx_rate.txt
for pz in 2 1 2 1; do MVB_POLY=$pz python tools/gpu_bench_attention.py --variants 0 --levels 0 --iters 20 2>&1 | grep ATTN_BENCH; done | tee $o/${tag}_attn_poly1.txt
