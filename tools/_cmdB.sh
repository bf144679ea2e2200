set -u
mkdir -p gpurun_out
for v in 1 0; do
  FISHING_K2_FOLD=$v timeout 300 python tools/k2_bench.py 2> gpurun_out/fold_k2_$v.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['v1_escapement_sigma0.01_autoreset']; print('fold=$v', 'auto-reset %.4g'%a['env_steps_per_s'], a['ret_hash'], 'plain %.4g'%d['v1_escapement_sigma0.01']['env_steps_per_s'])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02u.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r02u.log
