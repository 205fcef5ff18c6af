mkdir -p gpurun_out
timeout 600 python bench.py --images 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2u.json 2> gpurun_out/b_r2u.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/b_r2u.json').read().strip().splitlines()[-1])
    print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['e2e_rgb'])
except Exception as e:
    print('ERR',e, open('gpurun_out/b_r2u.err').read()[-1500:])
PY
