mkdir -p gpurun_out
timeout 600 python bench.py --images 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2t.json 2> gpurun_out/b_r2t.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/b_r2t.json').read().strip().splitlines()[-1])
    print('value %.1f'%d['value'], json.dumps(d['match'])[:1500])
except Exception as e:
    print('ERR',e, open('gpurun_out/b_r2t.err').read()[-1500:])
PY
