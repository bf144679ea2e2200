set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r02r.log
python tools/k2_bench.py > gpurun_out/k2_r02r.json 2>gpurun_out/k2_r02r.err; echo "k2 rc=$?"
( time python bench.py > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err ) 2>&1 | grep real; echo "bench rc=$?"
( time python bench.py --impl reference > gpurun_out/bench_ref_r02r.json 2> gpurun_out/bench_ref_r02r.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k2_r02r.json").read().strip().splitlines()[-1])
print({k: "%.4g"%v["env_steps_per_s"] for k,v in d.items() if isinstance(v,dict) and "env_steps_per_s" in v})
b=json.loads(open("gpurun_out/bench_r02r.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f frac %.3f e2e %.4g"%(b["value"], b["ms_per_step"], b["roofline"]["frac"], b["e2e"]["value"]))
print({k: "%.4g"%v["value"] for k,v in b["fused"].items()}, {k: "%.4g"%v["value"] for k,v in b["configs"].items()})
r=json.loads(open("gpurun_out/bench_ref_r02r.json").read().strip().splitlines()[-1]); print("ref %.4g"%r["value"], r["cpu_baseline"]["kind"])
PY
