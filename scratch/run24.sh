mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
timeout 600 $T --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/b_r2v_n$n.json 2> gpurun_out/b_r2v_n$n.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2v_n1.json 2> gpurun_out/b_r2v_n1.err
for n in 1 2 4 8; do python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2v_n'+n+'.json').read().strip().splitlines()[-1])
    print('N=%s value %.1f e2e %.1f'%(n, d['value'],d['e2e']['value']), d['collection']['equals_one_gpu_run'], round(d['collection']['ms'],2))
except Exception as e:
    print(n,'ERR',e, open('gpurun_out/b_r2v_n'+n+'.err').read()[-800:])
PY
done
