#!/bin/bash
# One profiling pass on the GPU box (run under gpurun, ONE GPU): every command runs plain first and must exit 0, then
# under ncu.  Outputs land in gpurun_out/ with the tag given as $1 (e.g. r01d); tools/ncu_summary.py turns the .ncu-rep
# files into the markdown tables kept under profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --T 50 --no-cpu-baseline --no-configs --e2e-T 16 --e2e-chunk 8"
NCU="ncu --clock-control none"

$BENCH > $OUT/plain_bench_$TAG.json 2> $OUT/plain_bench_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/plain_bench_$TAG.err; exit 1; }
# launch list of the bench command
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
# K1 headline kernel, warm (skip the first launches), full set
$NCU --set full -k regex:step_kernel_vec2 --launch-skip 40 --launch-count 3 -o $OUT/k1_$TAG -f $BENCH > $OUT/ncu_k1_$TAG.log 2>&1
# K2 fast path and the sweep kernel
python tools/k2_bench.py > $OUT/plain_k2_$TAG.json 2> $OUT/plain_k2_$TAG.err || { echo "plain k2 failed"; exit 1; }
K2B_T=200 $NCU --set full -k regex:rollout_fast_kernel --launch-skip 2 --launch-count 2 -o $OUT/k2_$TAG -f python tools/k2_bench.py > $OUT/ncu_k2_$TAG.log 2>&1
K2B_T=200 $NCU --set full -k regex:rollout_sweep --launch-skip 2 --launch-count 1 -o $OUT/k2sweep_$TAG -f python tools/k2_bench.py > $OUT/ncu_k2s_$TAG.log 2>&1
# next row f1: the K1 kernel of fishing-v8 (Myers: two logs, a power and an exp per step, all table-assisted) at 2^20 envs
if [ "${PROFILE_F1:-1}" != 0 ]; then
  python benchmarks/run_configs.py --only f1 > $OUT/plain_f1_$TAG.jsonl 2> $OUT/plain_f1_$TAG.err || { echo "plain f1 failed"; exit 1; }
  $NCU --set full --kernel-name-base demangled -k "regex:step_kernel_vec2<\\(int\\)4," --launch-skip 60 --launch-count 1 -o $OUT/k1v8_$TAG -f python benchmarks/run_configs.py --only f1 > $OUT/ncu_k1v8_$TAG.log 2>&1
fi
summarise() {   # the tables are made on the box: only text has to travel back (gpurun merges at most 64 MiB)
  python tools/launch_table.py $OUT/launches_$TAG.csv > $OUT/launch_table_$TAG.md 2>&1
  python tools/ncu_summary.py $OUT/*_$TAG.ncu-rep > $OUT/ncu_tables_$TAG.md 2>&1
  if [ "$(du -sm $OUT | cut -f1)" -gt 48 ]; then rm -f $OUT/*_$TAG.ncu-rep; fi
  ls -la $OUT/*_$TAG*
}
# fishing-v4 K1 at cfg-4 size (DRAM-resident state); PROFILE_CFG4=0 skips it
[ "${PROFILE_CFG4:-1}" = 0 ] && { summarise; exit 0; }
python benchmarks/run_configs.py --only cfg4 > $OUT/plain_cfg4_$TAG.json 2> $OUT/plain_cfg4_$TAG.err || { echo "plain cfg4 failed"; exit 1; }
$NCU --set full -k regex:step_kernel_vec2 --launch-skip 300 --launch-count 2 -o $OUT/k1v4_$TAG -f python benchmarks/run_configs.py --only cfg4 > $OUT/ncu_k1v4_$TAG.log 2>&1
summarise
