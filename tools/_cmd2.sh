set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "nccl or two_ranks" > gpurun_out/pytest_2gpu_r02s.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_2gpu_r02s.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_r02s.json 2> gpurun_out/bench_2gpu_r02s.err; echo "bench rc=$?"
python -c "
import json
b=json.loads(open('gpurun_out/bench_2gpu_r02s.json').read().strip().splitlines()[-1]); print('value %.4g ms %.3f e2e %.4g'%(b['value'], b['ms_per_step'], b['e2e']['value']), {k: ('%.4g'%v['value'], v.get('checksum')) for k,v in b['configs'].items()})"
