mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rf -s -p no:cacheprovider > gpurun_out/pytest_multi_r2x.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi_r2x.log
tail -5 gpurun_out/pytest_multi_r2x.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/b_r2x_n2.json 2> gpurun_out/b_r2x_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/b_r2x_n2.json').read().strip().splitlines()[-1])
    print('n2 value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), d.get('collection'))
except Exception as e: print('ERR',e,open('gpurun_out/b_r2x_n2.err').read()[-500:])
PY
