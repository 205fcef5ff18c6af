mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -k "write_exactly or graph" > gpurun_out/pytest_r2k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2k.log
tail -8 gpurun_out/pytest_r2k.log | cut -c1-300
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/b_r2k_n8.json 2> gpurun_out/b_r2k_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/b_r2k_n8.json').read().strip().splitlines()[-1])
    print('N=8 value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), d['collection'])
except Exception as e:
    print('ERR', e, open('gpurun_out/b_r2k_n8.err').read()[-1500:])
PY
timeout 600 $T --master-port 29512 bench.py --gpus 8 --workload collection --sets 512 --per-set 20000 --steps 2 --warmup 1 > gpurun_out/b_r2k_coll_n8.json 2> gpurun_out/b_r2k_coll_n8.err
cat gpurun_out/b_r2k_coll_n8.json; tail -n 3 gpurun_out/b_r2k_coll_n8.err
