#!/bin/bash
# usage: tools/ab_run.sh <tag> [variant ...]   -> gpurun_out/ab_<tag>.jsonl (one k2_bench line per build; "default" = the product)
tag=$1; shift
out=gpurun_out/ab_${tag}.jsonl
mkdir -p gpurun_out
: > $out
for v in default "$@"; do
  if [ "$v" = default ]; then lib=""; else lib=$PWD/gym_fishing_b200/ab/lib_${v}.so; fi
  FISHING_B200_LIB=$lib timeout 300 python tools/k2_bench.py >> $out 2>> gpurun_out/ab_${tag}.err || echo "{\"lib\": \"$v\", \"failed\": true}" >> $out
done
cat $out
