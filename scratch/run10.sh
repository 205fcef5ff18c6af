mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_r2i.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2i.log
tail -5 gpurun_out/pytest_r2i.log
B="timeout 300 python bench.py --images 64 --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/b_r2i_default.json 2> gpurun_out/b_r2i_default.err
for v in c4 c4b4 c2b6; do
SIFT_B200_LIB=$PWD/scratch/variants/libsift_b200_$v.so $B > gpurun_out/b_r2i_$v.json 2> gpurun_out/b_r2i_$v.err
done
for f in default c4 c4b4 c2b6; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2i_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), 'describe %.3f'%d['stages_ms']['describe'], 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2i_'+f+'.err').read()[-600:])
PY
done
