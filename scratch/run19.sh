mkdir -p gpurun_out
B="timeout 300 python bench.py --images 64 --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/b_r2q_diet.json 2> gpurun_out/b_r2q_diet.err
SIFT_B200_LIB=$PWD/scratch/variants/libsift_b200_nodiet.so $B > gpurun_out/b_r2q_nodiet.json 2> gpurun_out/b_r2q_nodiet.err
for v in diet nodiet; do
python - $v <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2q_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f'%d['value'], 'describe %.3f'%(d['stages_ms']['describe']), 'lat %.3f'%d['latency']['ms_per_image_one_stream'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2q_'+f+'.err').read()[-300:])
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --maxfail=5 -p no:cacheprovider -k "descriptor or golden or config1 or reproducible or 4k_set or 8k_set" > gpurun_out/pytest_r2q.log 2>&1
tail -3 gpurun_out/pytest_r2q.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/parity_report.json'))
for k in ('desc_stage_256x192','desc_stage_400x300','desc_stage_640x480','config3_4k_set','config4_8k_set'):
    v=r.get(k); print(k, json.dumps(v.get('desc_stage', v), default=float)[:200], (v.get('desc') or {}).get('frac_exact'))
PY
