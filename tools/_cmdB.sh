set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "v4 or cfg4" > gpurun_out/pytest_r02p.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r02p.log
: > gpurun_out/cfg4_r02p.jsonl
for v in default nostream v4s8 v4s9; do
  lib=""; st=1
  if [ "$v" = nostream ]; then st=0; elif [ "$v" != default ]; then lib=$PWD/gym_fishing_b200/ab/lib_${v}.so; fi
  FISHING_V4_STREAM=$st FISHING_B200_LIB=$lib timeout 300 python benchmarks/run_configs.py --only cfg4 2>> gpurun_out/cfg4_r02p.err | sed "s/^{/{\"lib\": \"$v\", /" >> gpurun_out/cfg4_r02p.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/cfg4_r02p.jsonl'):
    d=json.loads(l); print(d['lib'], 'py %.4g'%d['env_steps_per_s_env_only'], 'c_loop %.4g'%d['env_steps_per_s_env_only_c_loop'], 'gentle %.4g'%d['env_steps_per_s_env_only_c_loop_gentle_actions'])
PY
tail -3 gpurun_out/cfg4_r02p.err
