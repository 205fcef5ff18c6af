mkdir -p gpurun_out
B="timeout 300 python bench.py --images 64 --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/b_r2n_c2p8.json 2> gpurun_out/b_r2n_c2p8.err
for v in c2p0 c4p8 c4p8b4; do
SIFT_B200_LIB=$PWD/scratch/variants/libsift_b200_$v.so $B > gpurun_out/b_r2n_$v.json 2> gpurun_out/b_r2n_$v.err
done
for f in c2p8 c2p0 c4p8 c4p8b4; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2n_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), 'describe %.3f orient %.3f'%(d['stages_ms']['describe'],d['stages_ms']['orient']), 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'], 'kp', d['config']['keypoints_per_image'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2n_'+f+'.err').read()[-600:])
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --maxfail=5 -p no:cacheprovider -k "descriptor or golden or config1 or reproducible or 4k_set" > gpurun_out/pytest_r2n.log 2>&1
tail -3 gpurun_out/pytest_r2n.log
