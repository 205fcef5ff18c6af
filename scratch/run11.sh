mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rf -s -p no:cacheprovider > gpurun_out/pytest_r2j_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2j_multi.log
grep -E "passed|failed|collection ok|batch:|differ|FAILED" gpurun_out/pytest_r2j_multi.log | tail -12
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_r2j_n2.json 2> gpurun_out/b_r2j_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_r2j_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), d['collection'])
PY
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_r2j_n1.json 2> gpurun_out/b_r2j_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_r2j_n1.json').read().strip().splitlines()[-1])
print('N=1 value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), 'lat', d['latency']['ms_per_image_one_stream'], d['clocks'], {k:round(v,3) for k,v in d['stages_ms'].items()})
PY
