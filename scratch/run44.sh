mkdir -p gpurun_out
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/b_r2z_n$N.json 2> gpurun_out/b_r2z_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2z_n%s.json'%N).read().strip().splitlines()[-1])
    print('N',N,'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), d.get('collection'), d['clocks'])
except Exception as e: print('ERR',e,open('gpurun_out/b_r2z_n%s.err'%N).read()[-500:])
PY
