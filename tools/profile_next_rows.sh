#!/bin/bash
# Profiling pass for the rows beside the headline kernel (run under gpurun, ONE GPU): the lock-step kernel of the log-normal
# family (fishing-v5 Allen, fishing-v8 Myers) at 2^20 envs and of fishing-v4 at cfg-4 size.  Each command runs plain first and
# must exit 0, then under ncu --set full; tools/ncu_summary.py makes the tables on the box (only text travels back).
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
python benchmarks/run_configs.py --only f1 > $OUT/plain_f1_$TAG.jsonl 2> $OUT/plain_f1_$TAG.err || { echo "plain f1 failed"; tail -5 $OUT/plain_f1_$TAG.err; exit 1; }
$NCU --set full --kernel-name-base demangled -k "regex:step_kernel_vec2<\\(int\\)2," --launch-skip 60 --launch-count 1 -o $OUT/k1v5_$TAG -f python benchmarks/run_configs.py --only f1 > $OUT/ncu_k1v5_$TAG.log 2>&1
$NCU --set full --kernel-name-base demangled -k "regex:step_kernel_vec2<\\(int\\)4," --launch-skip 60 --launch-count 1 -o $OUT/k1v8_$TAG -f python benchmarks/run_configs.py --only f1 > $OUT/ncu_k1v8_$TAG.log 2>&1
python benchmarks/run_configs.py --only cfg4 > $OUT/plain_cfg4_$TAG.json 2> $OUT/plain_cfg4_$TAG.err || { echo "plain cfg4 failed"; exit 1; }
$NCU --set full -k regex:step_kernel_vec2 --launch-skip 300 --launch-count 2 -o $OUT/k1v4_$TAG -f python benchmarks/run_configs.py --only cfg4 > $OUT/ncu_k1v4_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/*_$TAG.ncu-rep > $OUT/ncu_tables_$TAG.md 2>&1
rm -f $OUT/*_$TAG.ncu-rep
ls -la $OUT/*_$TAG*
